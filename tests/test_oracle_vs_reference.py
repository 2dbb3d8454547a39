"""CPU, build container only: pin the oracle restatements to the reference's own modules imported
from /root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import geometries, ref_shim, tte_oracle, voc_oracle
from parrot_tts_b200 import configs, synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.mark.parametrize("seed,B,S", [(1, 1, 9), (2, 2, 24)])
def test_tte_oracle_vs_reference(seed, B, S):
    cfg = configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=seed)
    ref = ref_shim.load_reference_parrot(cfg, 150, 0, 10)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    phones = torch.randint(2, 150, (B, S), generator=g)
    if B > 1:
        phones[1, S // 2:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.arange(B) % 10}
    with torch.no_grad():
        logits, _, tgt_mask, log_dur = ref(batch, inference=True)
        ids_ref = ref.infer(batch)
    ids, r = tte_oracle.parrot_infer(sd, batch)
    assert ids == ids_ref
    assert torch.equal(r["tgt_mask"], tgt_mask)
    np.testing.assert_allclose(r["log_dur"].numpy(), log_dur.numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(r["logits"].numpy(), logits.numpy(), atol=5e-4, rtol=0)


def test_tte_reference_rejects_overlong():
    # reference modules/fft.py:18: pe[L] with L >= max_len raises IndexError
    cfg = configs.default_tte_config(max_len=16)
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0)
    phones = torch.randint(2, 150, (1, 16))
    with pytest.raises(IndexError):
        tte_oracle.parrot_forward(sd, {"phones": phones, "src_mask": phones != 0, "speaker": torch.zeros(1, dtype=torch.long)})


@pytest.mark.parametrize("seed,B,T", [(3, 1, 7), (4, 2, 33)])
def test_voc_oracle_vs_reference(seed, B, T):
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=seed, weight_norm=True)
    gen = ref_shim.load_reference_vocoder(h)
    gen.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    code = torch.randint(0, 1000, (B, T), generator=g)
    spkr = torch.randint(0, 10, (B, 1), generator=g)
    with torch.no_grad():
        ref = gen(code=code, spkr=spkr)
    wav = voc_oracle.code_generator_forward(sd, h, code, spkr)
    np.testing.assert_allclose(wav.numpy(), ref.numpy(), atol=2e-5, rtol=0)
    gen.remove_weight_norm()
    sd_plain = {k: v for k, v in gen.state_dict().items()}
    assert "conv_pre.weight" in sd_plain
    wav2 = voc_oracle.code_generator_forward(sd_plain, h, code, spkr)
    np.testing.assert_allclose(wav2.numpy(), ref.numpy(), atol=2e-5, rtol=0)


def test_voc_oracle_vs_reference_other_geometry():
    """The HiFi-GAN shape of tests/test_voc_gpu.py::test_other_hifigan_geometry_matches_reference_golden (x16, kernels 5 / 9, dilations (1, 2, 4) /
    (1, 3, 5)): the reference builds it from the same `h` (utils/vocoder/models.py:72-93) - pins the oracle there too."""
    h = geometries.other_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=23, weight_norm=True)
    gen = ref_shim.load_reference_vocoder(h)
    gen.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(77)
    code = torch.randint(0, 1000, (3, 300), generator=g)
    spkr = torch.randint(0, 10, (3, 1), generator=g)
    with torch.no_grad():
        ref = gen(code=code, spkr=spkr)
    wav = voc_oracle.code_generator_forward(sd, h, code, spkr)
    assert ref.shape == (3, 1, 4800)
    np.testing.assert_allclose(wav.numpy(), ref.numpy(), atol=2e-5, rtol=0)


@pytest.mark.parametrize("n_head", [2, 4])
def test_tte_oracle_vs_reference_other_geometry(n_head):
    """The transformer shape of tests/test_tte_gpu.py::test_other_transformer_geometry_matches_reference_golden (2 / 3 layers, FFN 512 x 5 taps,
    500 codes, 2 or 4 heads), built by the reference from the same config (modules/parrot.py:22-53)."""
    cfg = geometries.other_tte_config(n_head=n_head)
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=31)
    ref = ref_shim.load_reference_parrot(cfg, 150, 0, 10)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(12)
    phones = torch.randint(2, 150, (3, 41), generator=g)
    phones[1, 17:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([1, 7, 3])}
    with torch.no_grad():
        logits, _, tgt_mask, log_dur = ref(batch, inference=True)
        ids_ref = ref.infer(batch)
    ids, r = tte_oracle.parrot_infer(sd, batch, n_head=n_head)
    assert ids == ids_ref and logits.shape[-1] == 500
    assert torch.equal(r["tgt_mask"], tgt_mask)
    np.testing.assert_allclose(r["log_dur"].numpy(), log_dur.numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(r["logits"].numpy(), logits.numpy(), atol=5e-4, rtol=0)


def test_tte_oracle_vs_reference_single_speaker():
    # speakers.json with one entry: the reference builds no speaker_emb (modules/parrot.py:28-32) and adds nothing (:98-99)
    cfg = configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 1, seed=5)
    assert "speaker_emb.weight" not in sd
    ref = ref_shim.load_reference_parrot(cfg, 150, 0, 1)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    phones = torch.randint(2, 150, (2, 11), generator=g)
    phones[0, 7:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.zeros(2, dtype=torch.long)}
    with torch.no_grad():
        ids_ref = ref.infer(batch)
    ids, _ = tte_oracle.parrot_infer(sd, batch)
    assert ids == ids_ref


def test_tte_oracle_vs_reference_teacher_forced_and_zero_durations():
    # Parrot.forward(batch) with batch['duration'] (modules/parrot.py:101-103): includes zero-length characters and an item
    # whose total is far below the batch maximum (rows of zeros + pe row, still live keys up to len + 1)
    cfg = configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=6)
    ref = ref_shim.load_reference_parrot(cfg, 150, 0, 10)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(6)
    phones = torch.randint(2, 150, (3, 8), generator=g)
    phones[2, 3:] = 0
    dur = torch.randint(0, 5, (3, 8), generator=g)
    dur[1] = 0
    dur[1, 4] = 2
    dur[2, 3:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([1, 0, 9]), "duration": dur,
             "tgt_mask": None}              # the reference indexes batch['tgt_mask'] on this path; None = derive it from the durations
    with torch.no_grad():
        logits, _, tgt_mask, log_dur = ref(batch)
    r = tte_oracle.parrot_forward(sd, batch, dur_override=dur)
    assert torch.equal(r["tgt_mask"], tgt_mask) and r["logits"].shape == logits.shape
    assert [int(n) for n in r["out_lens"]] == dur.sum(1).tolist()
    np.testing.assert_allclose(r["logits"].numpy(), logits.numpy(), atol=5e-4, rtol=0)
    np.testing.assert_allclose(r["log_dur"].numpy(), log_dur.numpy(), atol=2e-5, rtol=0)


def test_tte_oracle_vs_reference_supplied_tgt_mask():
    # the training-side batch carries tgt_mask = codes != pad (exact lengths, modules/data.py:117-119); the reference's
    # length_regulator returns it unchanged and the decoder attends with it (modules/duration.py:9-10,22-24; parrot.py:104,107) -
    # one live key fewer per shorter item than the derived `<=` mask
    cfg = configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=16)
    ref = ref_shim.load_reference_parrot(cfg, 150, 0, 10)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(16)
    phones = torch.randint(2, 150, (3, 9), generator=g)
    phones[1, 5:] = 0
    dur = torch.randint(1, 5, (3, 9), generator=g)
    dur[1, 5:] = 0
    dur[2, :] = 1
    lens = dur.sum(1)
    T = int(lens.max())
    tgt_mask = torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)           # exact lengths, unlike get_mask_from_lengths
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([3, 0, 7]), "duration": dur, "tgt_mask": tgt_mask}
    with torch.no_grad():
        logits, _, mask_out, log_dur = ref(batch)
    assert mask_out is tgt_mask
    r = tte_oracle.parrot_forward(sd, batch, dur_override=dur, tgt_mask=tgt_mask)
    assert torch.equal(r["tgt_mask"], tgt_mask)
    np.testing.assert_allclose(r["logits"].numpy(), logits.numpy(), atol=5e-4, rtol=0)
    derived = tte_oracle.parrot_forward(sd, batch, dur_override=dur)        # the `<=` mask gives different logits for the shorter items
    assert float((derived["logits"][1] - r["logits"][1]).abs().max()) > 1e-3


def test_generator_oracle_vs_reference_feature_input():
    # Generator.forward on features (utils/vocoder/models.py:95-111) through the reference CodeGenerator's parent class
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=8, weight_norm=True)
    gen = ref_shim.load_reference_vocoder(h)
    gen.load_state_dict(sd, strict=True)
    x = torch.randn(2, h.model_in_dim, 9, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        ref = type(gen).__mro__[1].forward(gen, x)
    wav = voc_oracle.generator_forward(sd, h, x)
    np.testing.assert_allclose(wav.numpy(), ref.numpy(), atol=2e-5, rtol=0)
    # the reference's own int16 expression (utils/vocoder/inference.py:71-73), including a sample that wraps
    y = torch.tensor([[0.5, -0.25, 0.99999, -1.0, 1.0, 3.0517578125e-05 * 1.5]])
    want = (y.squeeze() * 32768.0).cpu().numpy().astype("int16")
    assert np.array_equal(voc_oracle.to_int16(y), want)


def test_checkpoint_layouts_are_interchangeable_with_the_reference():
    """Drop-in boundary: the mirror classes' state_dict() has exactly the reference modules' keys and
    shapes, each loads the other's with strict=True - with weight norm attached and after remove_weight_norm()."""
    from parrot_tts_b200.tte import Parrot
    from parrot_tts_b200.vocoder import CodeGenerator
    import json, os, tempfile
    d = tempfile.mkdtemp(prefix="parrot_ckpt_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    cfg = configs.default_tte_config(root_path=d)
    ours_t, ref_t = Parrot(cfg, 150, 0), ref_shim.load_reference_parrot(cfg, 150, 0, 10)
    a, b = ours_t.state_dict(), ref_t.state_dict()
    assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
    ours_t.load_state_dict(b, strict=True)
    ref_t.load_state_dict(a, strict=True)

    h = configs.default_vocoder_config()
    ours_v, ref_v = CodeGenerator(h), ref_shim.load_reference_vocoder(h)
    a, b = ours_v.state_dict(), ref_v.state_dict()
    assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
    assert "conv_pre.weight_g" in a and "conv_pre.weight_v" in a
    ours_v.load_state_dict(b, strict=True)
    ref_v.load_state_dict(a, strict=True)
    ours_v.remove_weight_norm()
    ref_v.remove_weight_norm()
    a, b = ours_v.state_dict(), ref_v.state_dict()
    assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
    assert "conv_pre.weight" in a and "conv_pre.weight_g" not in a
    ours_v.load_state_dict(b, strict=True)
    for k in b:                                      # and the folded weights really are the reference's g * v / ||v||
        assert torch.equal(ours_v.state_dict()[k], b[k])


def test_mel_oracle_vs_reference_function():
    # the reference's own mel_spectrogram (utils/vocoder/dataset.py:43-69), imported with its absent top-level imports stubbed
    from oracle import mel_oracle
    ref = ref_shim.load_reference_mel_spectrogram(
        lambda sr, n_fft, n_mels, fmin, fmax: mel_oracle.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax))
    y = (torch.rand(3, 5000, generator=torch.Generator().manual_seed(3)) * 2 - 1) * 0.9
    a = ref(y, 1024, 80, 16000, 256, 1024, 0, 8000)
    b = mel_oracle.mel_spectrogram(y, 1024, 80, 16000, 256, 1024, 0, 8000)
    assert a.shape == b.shape == (3, 80, 19)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=1e-6)
