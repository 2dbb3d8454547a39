"""ORACLE support (test infrastructure): generate tests/golden/*.npz by running the REFERENCE's own
modules (imported from /root/reference, see ref_shim.py) on CPU fp32 with the seeded synthetic
checkpoints of parrot_tts_b200/synth.py.

Run in the build container:   python oracle/gen_golden.py
The fixtures are committed; the GPU box (no /root/reference) only reads them.
Each file records the seed, torch version and the config so it can be regenerated.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.geometries import other_tte_config, other_vocoder_config  # noqa: E402
from parrot_tts_b200 import configs, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
VOCAB, PAD, NSPK = 150, 0, 10


def tte_case(name, phones, speaker, seed=0, dur_mode="spread", stride=5, cfg=None):
    cfg = cfg or configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, VOCAB, PAD, NSPK, seed=seed, dur_mode=dur_mode)
    model = ref_shim.load_reference_parrot(cfg, VOCAB, PAD, NSPK)
    model.load_state_dict(sd, strict=True)
    phones = torch.as_tensor(phones, dtype=torch.long)
    batch = {"phones": phones, "src_mask": phones != PAD, "speaker": torch.as_tensor(speaker, dtype=torch.long)}
    with torch.no_grad():
        logits, src_mask, tgt_mask, log_dur = model(batch, inference=True)
        ids_list = model.infer(batch)
    dur = torch.clamp(torch.round(torch.exp(log_dur) - 1), min=0).long()
    top2 = torch.topk(logits, 2, dim=-1).values
    ids_flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in ids_list])
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        phones=phones.numpy(), speaker=batch["speaker"].numpy(), src_mask=src_mask.numpy(),
        log_dur=log_dur.numpy(), dur=dur.numpy(), tgt_mask=tgt_mask.numpy(),
        ids_flat=ids_flat, ids_lens=np.asarray([len(x) for x in ids_list], dtype=np.int64),
        argmax=torch.argmax(logits, -1).numpy(),
        margin=(top2[..., 0] - top2[..., 1]).numpy(),
        logits_rows=logits[:, ::stride, :].numpy().astype(np.float32),
        logits_row_stride=np.int64(stride), dur_mode=str(dur_mode),
        seed=np.int64(seed), torch_version=str(torch.__version__), vocab=np.int64(VOCAB), n_speaker=np.int64(NSPK),
    )
    print(name, "T =", logits.shape[1], "lens", [len(x) for x in ids_list], "min margin", float((top2[..., 0] - top2[..., 1]).min()))


def voc_case(name, code, spkr, seed=0, x=None, h=None):
    h = h or configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=seed, weight_norm=True, n_speaker=NSPK)
    gen = ref_shim.load_reference_vocoder(h)
    gen.load_state_dict(sd, strict=True)
    out = {}
    with torch.no_grad():
        if x is None:
            code = torch.as_tensor(code, dtype=torch.long)
            spkr = torch.as_tensor(spkr, dtype=torch.long).view(-1, 1)
            wav = gen(code=code, spkr=spkr)
            out.update(code=code.numpy(), spkr=spkr.numpy())
        else:
            from importlib import import_module  # Generator.forward on a float feature map
            wav = type(gen).__mro__[1].forward(gen, x)
            out.update(x=x.numpy())
    out.update(wav=wav.numpy().astype(np.float32), seed=np.int64(seed), torch_version=str(torch.__version__))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, tuple(wav.shape), "abs max", float(wav.abs().max()), "rms", float(wav.pow(2).mean().sqrt()))


def mel_case(name, seed=0):
    """The reference's own mel_spectrogram (utils/vocoder/dataset.py:43-69) at the shipped config (n_fft 1024, hop 256, 80 mels,
    16 kHz, 0..8000 Hz) on noise + a chirp; the librosa mel basis (absent here) is the oracle's restatement - see ref_shim."""
    from oracle import mel_oracle
    ref = ref_shim.load_reference_mel_spectrogram(
        lambda sr, n_fft, n_mels, fmin, fmax: mel_oracle.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax))
    g = torch.Generator().manual_seed(seed)
    n = 16384
    t = torch.arange(n, dtype=torch.float64) / 16000.0
    chirp = torch.sin(2 * torch.pi * (200.0 * t + 1500.0 * t * t)).float()
    y = torch.stack([0.5 * chirp + 0.05 * torch.randn(n, generator=g), 0.8 * (torch.rand(n, generator=g) * 2 - 1)])
    y[1, 9000:] = 0.0                                     # a silent tail: exercises the clamp at 1e-5
    with torch.no_grad():
        mel = ref(y, 1024, 80, 16000, 256, 1024, 0, 8000)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), y=y.numpy(), mel=mel.numpy().astype(np.float32), seed=np.int64(seed),
                        torch_version=str(torch.__version__))
    print(name, tuple(mel.shape), "range", float(mel.min()), float(mel.max()))


def other_geometries():
    """Fixtures at the two other geometries (tests/test_*_gpu.py::test_other_*_geometry_matches_reference_golden)."""
    g3 = torch.Generator().manual_seed(77)
    voc_case("voc_geom_x16_t300", torch.randint(0, 1000, (3, 300), generator=g3), torch.randint(0, 10, (3,), generator=g3), seed=23,
             h=other_vocoder_config())
    g4 = torch.Generator().manual_seed(12)
    ph = torch.randint(2, VOCAB, (3, 41), generator=g4)
    ph[1, 17:] = PAD
    tte_case("tte_geom_l23_s41", ph, [1, 7, 3], seed=31, stride=16, cfg=other_tte_config())


def main():
    assert ref_shim.available(), "needs /root/reference (build container only)"
    os.makedirs(GOLDEN, exist_ok=True)
    if "--other-geometries" in sys.argv:      # only the fixtures added last (the others are not rewritten)
        return other_geometries()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)
    # TTE: one utterance; a ragged padded batch (rectangle semantics, SURVEY.md hard part 3)
    tte_case("tte_b1_s16", torch.randint(2, VOCAB, (1, 16), generator=g), [0])
    ph = torch.randint(2, VOCAB, (3, 20), generator=g)
    ph[1, 13:] = PAD
    ph[2, 7:] = PAD
    tte_case("tte_b3_ragged", ph, [0, 3, 7])
    # vocoder: code path B=1 and B=2, plus Generator.forward on a raw feature map
    voc_case("voc_b1_t24", torch.randint(0, 1000, (1, 24), generator=g), [2])
    voc_case("voc_b2_t40", torch.randint(0, 1000, (2, 40), generator=g), [0, 9])
    voc_case("gen_b1_t12", None, None, x=torch.randn(1, 256, 12, generator=g))
    # BASELINE configs[0] size, pinned to the REAL reference on CPU: 1 x 64 chars -> exactly 512 units
    # (the const8 recipe of SURVEY 8d) -> and a 512-unit utterance through the vocoder (163 840 samples)
    g2 = torch.Generator().manual_seed(4321)
    tte_case("tte_cfg1_s64", torch.randint(2, VOCAB, (1, 64), generator=g2), [0], dur_mode="const8", stride=64)
    voc_case("voc_cfg1_t512", torch.randint(0, 1000, (1, 512), generator=g2), [3])
    # a ragged batch at config-3 source length: 4 items, up to 256 chars, predicted ("spread") durations
    ph = torch.randint(2, VOCAB, (4, 256), generator=g2)
    ph[1, 200:] = PAD
    ph[2, 97:] = PAD
    ph[3, 31:] = PAD
    tte_case("tte_b4_s256_ragged", ph, [1, 8, 0, 5], seed=2, stride=128)
    mel_case("mel_b2_n16384")
    other_geometries()


if __name__ == "__main__":
    main()
