"""ORACLE (test infrastructure, not product code): CPU restatement of the reference HiFi-GAN
code generator (unit ids + speaker id -> waveform).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``parrot_tts_b200``) never does.

What it restates (citations relative to the reference tree):
  * weight_norm folding  w = g * v / ||v||, norm over dims != 0   utils/vocoder/models.py:17-28,75-93
                         (``torch._weight_norm(v, g, 0)`` - the op the reference's hook runs)
  * ResBlock1.forward                                            utils/vocoder/models.py:31-38
  * Generator.forward    (final leaky_relu uses slope 0.01)      utils/vocoder/models.py:95-111
  * CodeGenerator.forward + _upsample                            utils/vocoder/models.py:132-169
  * get_padding                                                  utils/vocoder/utils.py:44-45
  * generate(): x32768 -> int16                                  utils/vocoder/inference.py:65-74

Parity pin: checked against the reference's own ``CodeGenerator`` run in the build container
(tests/test_oracle_vs_reference.py) and against tests/golden/voc_*.npz generated from it by
oracle/gen_golden.py.  The reference ships no golden vectors of its own.
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1
MAX_WAV_VALUE = 32768.0


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def folded_weight(sd, key):
    """Effective conv weight for either checkpoint flavour (with / without weight norm)."""
    if key + ".weight" in sd:
        return sd[key + ".weight"]
    return torch._weight_norm(sd[key + ".weight_v"], sd[key + ".weight_g"], 0)


def resblock1(sd, p, x, k, dilations):
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, folded_weight(sd, f"{p}.convs1.{m}"), sd[f"{p}.convs1.{m}.bias"],
                      padding=get_padding(k, d), dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, folded_weight(sd, f"{p}.convs2.{m}"), sd[f"{p}.convs2.{m}.bias"],
                      padding=get_padding(k, 1), dilation=1)
        x = xt + x
    return x


def generator_forward(sd, h, x):
    """Generator.forward: f32 [B, model_in_dim, T] -> f32 [B, 1, prod(upsample_rates) * T]."""
    with torch.no_grad():
        nk = len(h["resblock_kernel_sizes"])
        x = F.conv1d(x, folded_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            x = F.leaky_relu(x, LRELU_SLOPE)
            x = F.conv_transpose1d(x, folded_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"],
                                   stride=u, padding=(k - u) // 2)
            xs = None
            for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
                r = resblock1(sd, f"resblocks.{i * nk + j}", x, rk, rd)
                xs = r if xs is None else xs + r
            x = xs / nk
        x = F.leaky_relu(x)  # default slope 0.01 (models.py:107)
        x = F.conv1d(x, folded_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
        return torch.tanh(x)


def code_generator_input(sd, h, code, spkr):
    """CodeGenerator.forward up to the Generator call: [B, 128+128, T]."""
    x = F.embedding(code, sd["dict.weight"]).transpose(1, 2)
    if h.get("multispkr", None):
        s = F.embedding(spkr, sd["spkr.weight"]).transpose(1, 2)  # [B, E, 1]
        s = s.unsqueeze(3).repeat(1, 1, 1, x.shape[-1] // s.shape[2]).view(s.shape[0], s.shape[1], x.shape[-1])
        x = torch.cat([x, s], dim=1)
    return x


def code_generator_forward(sd, h, code, spkr):
    with torch.no_grad():
        return generator_forward(sd, h, code_generator_input(sd, h, code, spkr))


def to_int16(y):
    # utils/vocoder/inference.py:71-73: (y * 32768).cpu().numpy().astype('int16')
    # numpy float->int16 truncates toward zero (C cast; wraps when |y*32768| >= 32768).
    return (y.squeeze() * MAX_WAV_VALUE).numpy().astype("int16")


def peak_normalize(audio):
    """`librosa.util.normalize(audio.astype(np.float32))` (utils/vocoder/inference.py:169, demo.ipynb cell 13).

    librosa is a third-party dependency that is NOT vendored in the reference (requirements.txt:8, unpinned) and is
    absent here, so this restates its published algorithm for the defaults the reference uses (norm=inf, axis=0,
    threshold=tiny, fill=None): magnitudes are taken in float64, the peak below `tiny` is replaced by 1, the quotient is
    stored into an array of the input's dtype (float32).  Parity for this one function is unpinned by any reference run."""
    import numpy as np
    s = np.asarray(audio).astype(np.float32)
    length = np.abs(s).astype(float).max(axis=0, keepdims=True)
    length[length < np.finfo(np.float32).tiny] = 1.0
    out = np.empty_like(s)
    out[:] = s / length
    return out

