"""Deterministic synthetic checkpoints in the reference's two state_dict layouts.

There is no network for the trained Google-Drive checkpoints (reference
demo.ipynb cell 4), so the parity tests, golden vectors and benchmarks all run
on seeded random weights generated here.  The tensors are a pure function of
(key name, seed), independent of any module construction order, so the very
same weights can be rebuilt on the GPU box where ``/root/reference`` does not
exist.

Layouts (names + shapes) follow SURVEY.md section 8b:
  * TTE:      keys of ``Parrot.state_dict()`` (reference modules/parrot.py:13-65)
  * vocoder:  ``checkpoint['generator']`` with old-style weight-norm keys
              ``weight_g``/``weight_v`` (reference utils/vocoder/models.py:17-28,75-93)
"""
import hashlib
import math

import torch


def _gen(key, seed):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def _randn(key, seed, *shape):
    return torch.randn(*shape, generator=_gen(key, seed), dtype=torch.float32)


def _uniform(key, seed, bound, *shape):
    return (torch.rand(*shape, generator=_gen(key, seed), dtype=torch.float32) * 2 - 1) * bound


def sinusoid_table(max_len, d_model):
    """The `pe` buffer (reference modules/fft.py:22-38): sin on even, cos on odd columns."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe


def synth_tte_state_dict(cfg, src_vocab_size=150, src_pad_idx=0, n_speaker=10, seed=0,
                         dur_mode="spread", head_gain=4.0):
    """Random TTE weights.

    dur_mode:
      "spread"  parity runs - default-scale duration projection with bias ln(6.5): durations
                spread over roughly 0..24 units per char.
      "const8"  throughput runs - proj.weight = 0, proj.bias = ln 9 -> every real char gets
                round(e^{ln 9} - 1) = 8 units (64 chars -> 512 units, 256 -> 2048).
    head_gain multiplies head.weight so the top-2 logit margins are not degenerate.
    """
    tr = cfg["transformer"]
    d = tr["d_model"]
    ff = tr["conv_n_filter"]
    k1, k2 = tr["conv_kernel_sizes"]
    dp = cfg["duration_predictor"]
    nf, dk = dp["n_filter"], dp["kernel_size"]
    n_codes = cfg["preprocess"]["hubert_codes"]
    sd = {}

    def lin(key, out_f, in_f, bias=True, gain=1.0):
        b = 1.0 / math.sqrt(in_f)
        sd[key + ".weight"] = _uniform(key + ".weight", seed, b * gain, out_f, in_f)
        if bias:
            sd[key + ".bias"] = _uniform(key + ".bias", seed, b, out_f)

    def conv(key, out_c, in_c, k):
        b = 1.0 / math.sqrt(in_c * k)
        sd[key + ".weight"] = _uniform(key + ".weight", seed, b, out_c, in_c, k)
        sd[key + ".bias"] = _uniform(key + ".bias", seed, b, out_c)

    def ln(key, n):
        sd[key + ".weight"] = 1.0 + 0.1 * _randn(key + ".weight", seed, n)
        sd[key + ".bias"] = 0.1 * _randn(key + ".bias", seed, n)

    sd["pos_emb.pe"] = sinusoid_table(tr["max_len"], d)
    tok = _randn("tok_emb.weight", seed, src_vocab_size, d)
    tok[src_pad_idx].zero_()
    sd["tok_emb.weight"] = tok
    if n_speaker > 1:
        sd["speaker_emb.weight"] = _randn("speaker_emb.weight", seed, n_speaker, d)

    conv("duration_predictor.layers.0.conv", nf, d, dk)
    ln("duration_predictor.layers.2", nf)
    conv("duration_predictor.layers.4.conv", nf, nf, dk)
    ln("duration_predictor.layers.6", nf)
    lin("duration_predictor.proj", 1, nf)
    if dur_mode == "const8":
        sd["duration_predictor.proj.weight"].zero_()
        sd["duration_predictor.proj.bias"].fill_(math.log(9.0))
    elif dur_mode == "spread":
        sd["duration_predictor.proj.bias"].fill_(math.log(6.5))
    else:
        raise ValueError(dur_mode)

    for stack, n_layer in (("encoder_layers", tr["encoder"]["n_layer"]),
                           ("decoder_layers", tr["decoder"]["n_layer"])):
        for i in range(n_layer):
            p = f"{stack}.{i}"
            lin(p + ".attention.qkv", 3 * d, d, bias=False)
            xb = math.sqrt(6.0 / (d + 3 * d))  # xavier_uniform bound of nn.MultiheadAttention
            sd[p + ".attention.mha.in_proj_weight"] = _uniform(p + ".attention.mha.in_proj_weight", seed, xb, 3 * d, d)
            lin(p + ".attention.mha.out_proj", d, d, bias=False)
            lin(p + ".attention.wo", d, d, bias=False)
            conv(p + ".convlayer.conv1", ff, d, k1)
            conv(p + ".convlayer.conv2", d, ff, k2)
            ln(p + ".attn_norm", d)
            ln(p + ".conv_norm", d)

    lin("head", n_codes, d, gain=head_gain)
    return sd


def synth_vocoder_state_dict(h, seed=0, weight_norm=True, n_speaker=10, out_gain=1.0):
    """Random CodeGenerator weights, pre-tanh amplitude O(1) so the 1e-3 abs tolerance is
    exercised at realistic (near full-scale) amplitude.

    weight_norm=True  -> old-style keys ``weight_g`` [C0,1,1] / ``weight_v`` (the layout of a
                         checkpoint saved before ``remove_weight_norm()``; norm over dims != 0,
                         which for ConvTranspose1d is per *input* channel).
    weight_norm=False -> plain ``weight`` keys holding g*v/||v||.
    """
    sd = {}

    def wn_conv(key, shape, fan_in, gain):
        v = _randn(key + ".weight_v", seed, *shape) * 0.01
        # per-dim-0 magnitude: He-style effective std, jittered so g carries information
        c0 = shape[0]
        n_per = 1
        for s in shape[1:]:
            n_per *= s
        target_std = gain * math.sqrt(2.0 / fan_in)
        jitter = 1.0 + 0.2 * _uniform(key + ".weight_g", seed, 1.0, c0)
        g = (target_std * math.sqrt(n_per) * jitter).reshape(c0, *([1] * (len(shape) - 1)))
        b = 0.05 * _randn(key + ".bias", seed, shape[0] if key.startswith(("conv", "res")) else shape[1])
        if weight_norm:
            sd[key + ".weight_g"] = g
            sd[key + ".weight_v"] = v
        else:
            sd[key + ".weight"] = torch._weight_norm(v, g, 0)
        sd[key + ".bias"] = b

    c0 = h["upsample_initial_channel"]
    in_dim = h.get("model_in_dim", 128)
    wn_conv("conv_pre", (c0, in_dim, 7), in_dim * 7, 1.0)
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        # each output sample sees cin * k / u taps
        wn_conv(f"ups.{i}", (cin, cout, k), cin * k / u, 1.0)
        ch = cout
    nk = len(h["resblock_kernel_sizes"])
    for i in range(len(h["upsample_rates"])):
        ch = c0 // (2 ** (i + 1))
        for j, (k, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            r = i * nk + j
            for m in range(len(dil)):
                wn_conv(f"resblocks.{r}.convs1.{m}", (ch, ch, k), ch * k, 1.0)
                wn_conv(f"resblocks.{r}.convs2.{m}", (ch, ch, k), ch * k, 0.4)
    wn_conv("conv_post", (1, ch, 7), ch * 7, 0.15 * out_gain)
    sd["dict.weight"] = _randn("dict.weight", seed, h["num_embeddings"], h["embedding_dim"])
    if h.get("multispkr", None):
        sd["spkr.weight"] = _randn("spkr.weight", seed, n_speaker, h["embedding_dim"])
    return sd
