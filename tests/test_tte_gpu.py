"""GPU parity: TTE kernels and the full Parrot path against the CPU oracle and the golden vectors recorded
from the reference.  Integer outputs (durations, unit ids, masks) must be bit-exact; floats within fp32
round-off (tolerances written per assert)."""
import json
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import geometries, tte_oracle
from parrot_tts_b200 import configs, ops, synth
from parrot_tts_b200.tte import DurationPredictor, FFTBlock, LitParrot, Parrot, SinusoidalPosEmb, length_regulator

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VOCAB, NSPK = 150, 10


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _cfg(max_len=None):
    d = tempfile.mkdtemp(prefix="parrot_tte_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(NSPK)}, f)
    return configs.default_tte_config(root_path=d, max_len=max_len)


def _parrot(seed=0, max_len=None, **synth_kw):
    cfg = _cfg(max_len)
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=seed, **synth_kw)
    m = Parrot(cfg, VOCAB, 0)
    m.load_state_dict(sd, strict=True)
    return cfg, sd, m.eval().to(DEV)


def _to_dev(batch):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.mark.parametrize("M,N,K", [(1, 256, 256), (130, 768, 256), (257, 1000, 256), (64, 256, 1024), (300, 100, 36)])
def test_linear_kernel(M, N, K):
    g = _gen(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = x @ w.t() + b
    got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())


@pytest.mark.parametrize("M,N,K", [(1, 256, 256), (130, 768, 256), (257, 1024, 256), (64, 256, 1024), (300, 112, 64), (129, 16, 32)])
def test_linear_kernel_tensor_core(M, N, K):
    """tcgen05 GEMM on split-fp16 operands (3 passes, fp32 accumulation): fp32-class accuracy."""
    g = _gen(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = (x.double() @ w.double().t() + b.double()).float()
    got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), mode=1).cpu()
    ref32 = x @ w.t() + b
    err, err32 = float((got - ref).abs().max()), float((ref32 - ref).abs().max())
    assert err < 2e-5 and err < 20 * max(err32, 1e-7), (err, err32)


def test_layernorm_kernel():
    g = _gen(1)
    x = torch.randn(37, 256, generator=g) * 3 + 1
    w, b = torch.randn(256, generator=g), torch.randn(256, generator=g)
    ref = F.layer_norm(x, (256,), w, b, 1e-5)
    got = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=5e-6, rtol=1e-5), float((got - ref).abs().max())


def test_length_regulator_matches_oracle():
    g = _gen(2)
    seq = torch.randn(3, 11, 256, generator=g)
    dur = torch.randint(0, 6, (3, 11), generator=g)
    dur[2] = 0
    dur[2, 3] = 2                       # an item much shorter than the batch maximum
    ref, ref_mask, _ = tte_oracle.length_regulator(seq, dur)
    got, mask = length_regulator(seq.to(DEV), dur.to(DEV))
    assert torch.equal(got.cpu(), ref)              # pure gather: bit-exact
    assert torch.equal(mask.cpu(), ref_mask)        # `<=` off-by-one mask (data.py:18)


def test_pos_emb_single_row_quirk():
    pe = SinusoidalPosEmb(64, 256).to(DEV)
    x = torch.randn(2, 9, 256, generator=_gen(3))
    got = pe(x.to(DEV)).cpu()
    assert torch.equal(got, pe.pe.cpu()[9] + x)     # row index = sequence LENGTH, same row for every position
    with pytest.raises(IndexError):
        pe(torch.zeros(1, 64, 256, device=DEV))


@pytest.mark.parametrize("fold,mode", [(True, 0), (False, 0), (True, 1)])
@pytest.mark.parametrize("B,L", [(1, 5), (2, 37), (3, 130)])
def test_fft_block_matches_oracle(fold, mode, B, L):
    cfg = _cfg()
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=5)
    p = "decoder_layers.1"
    blk = FFTBlock(256, 2, 1024, [9, 1], 0.1)
    blk.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")}, strict=True)
    blk = blk.eval().to(DEV)
    blk.fold_projections = fold
    blk.gemm_mode = mode
    g = _gen(B * 100 + L)
    x = torch.randn(B, L, 256, generator=g)
    pad = torch.zeros(B, L, dtype=torch.bool)
    if B > 1:
        pad[1, L // 2:] = True
    with torch.no_grad():
        ref = tte_oracle.fft_block(sd, p, x, pad, 2)
    got = blk(x.to(DEV), key_padding_mask=pad.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), float((got - ref).abs().max())


def test_duration_predictor_matches_oracle():
    cfg = _cfg()
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=6)
    dp = DurationPredictor(256, 256, 3, 0.5)
    dp.load_state_dict({k[len("duration_predictor."):]: v for k, v in sd.items() if k.startswith("duration_predictor.")}, strict=True)
    dp = dp.eval().to(DEV)
    g = _gen(7)
    x = torch.randn(2, 23, 256, generator=g)
    mask = torch.zeros(2, 23, dtype=torch.bool)
    mask[1, 15:] = True
    with torch.no_grad():
        ref = tte_oracle.duration_predictor(sd, x, mask)
    got = dp(x.to(DEV), mask.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())
    assert (got[1, 15:] == 0).all()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["tte_b1_s16", "tte_b3_ragged", "tte_cfg1_s64", "tte_b4_s256_ragged"])
def test_parrot_matches_reference_golden(golden_dir, name, mode):
    """Against vectors recorded from the reference's own modules on CPU (oracle/gen_golden.py), up to BASELINE configs[0] size
    (64 chars -> 512 units) and a ragged batch at config-3 source length: ZERO id mismatches, durations and masks bit-exact."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    _, _, m = _parrot(seed=int(g["seed"]), dur_mode=str(g["dur_mode"]))
    m.gemm_mode = mode
    batch = _to_dev({"phones": torch.from_numpy(g["phones"]), "src_mask": torch.from_numpy(g["src_mask"]),
                     "speaker": torch.from_numpy(g["speaker"])})
    logits, src_mask, tgt_mask, log_dur = m(batch, inference=True)
    np.testing.assert_allclose(log_dur.cpu().numpy(), g["log_dur"], rtol=0, atol=5e-5)
    assert np.array_equal(tgt_mask.cpu().numpy(), g["tgt_mask"])                   # implies bit-exact durations
    ids = m.infer(batch)
    assert [len(x) for x in ids] == g["ids_lens"].tolist()
    flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in ids])
    mism = np.nonzero(flat != g["ids_flat"])[0]
    am = torch.argmax(logits, -1).cpu().numpy()
    bad = np.argwhere(am != g["argmax"])
    detail = [(tuple(ix), float(g["margin"][tuple(ix)])) for ix in bad[:5]]
    assert mism.size == 0 and bad.size == 0, f"{mism.size} id mismatches; first (index, reference top-2 margin): {detail}"
    stride = int(g["logits_row_stride"])
    np.testing.assert_allclose(logits[:, ::stride].cpu().numpy(), g["logits_rows"], rtol=0, atol=2e-3)


@pytest.mark.parametrize("fold,mode", [(True, 0), (False, 0), (True, 1)])
def test_parrot_vs_oracle_random_batch(fold, mode):
    cfg, sd, m = _parrot(seed=21)
    m.fold_projections = fold
    m.gemm_mode = mode
    g = _gen(8)
    phones = torch.randint(2, VOCAB, (4, 33), generator=g)
    phones[1, 20:] = 0
    phones[3, 5:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([0, 9, 4, 2])}
    ids_ref, r = tte_oracle.parrot_infer(sd, batch)
    ids = m.infer(_to_dev(batch))
    logits, _, tgt_mask, log_dur = m(_to_dev(batch), inference=True)
    assert torch.equal(tgt_mask.cpu(), r["tgt_mask"])
    assert torch.allclose(log_dur.cpu(), r["log_dur"], atol=5e-5, rtol=0)
    top2 = torch.topk(r["logits"], 2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    n_bad = 0
    for b, (a, e) in enumerate(zip(ids, ids_ref)):
        assert len(a) == len(e)
        for t, (u, v) in enumerate(zip(a, e)):
            if u != v:
                n_bad += 1
                assert margin[b, t] < 1e-4, f"id mismatch at a clear-cut position b={b} t={t} margin={float(margin[b, t])}"
    assert n_bad == 0, f"{n_bad} near-tie id mismatches"
    assert torch.allclose(logits.cpu(), r["logits"], atol=3e-3, rtol=0), float((logits.cpu() - r["logits"]).abs().max())


def test_parrot_teacher_forced_durations():
    cfg, sd, m = _parrot(seed=22)
    g = _gen(9)
    phones = torch.randint(2, VOCAB, (2, 10), generator=g)
    dur = torch.randint(0, 5, (2, 10), generator=g)
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([1, 2]), "duration": dur}
    r = tte_oracle.parrot_forward(sd, batch, dur_override=dur)
    logits, _, tgt_mask, _ = m(_to_dev(batch), inference=False)
    assert torch.equal(tgt_mask.cpu(), r["tgt_mask"])
    assert torch.allclose(logits.cpu(), r["logits"], atol=3e-3, rtol=0)


def test_parrot_teacher_forced_with_supplied_tgt_mask():
    """batch['tgt_mask'] (exact lengths, as the reference's dataset builds it) is the decoder's key padding mask and is returned
    unchanged (reference modules/duration.py:9-10,22-24; parrot.py:104,107)."""
    cfg, sd, m = _parrot(seed=23)
    g = _gen(10)
    phones = torch.randint(2, VOCAB, (3, 12), generator=g)
    phones[1, 7:] = 0
    dur = torch.randint(1, 5, (3, 12), generator=g)
    dur[1, 7:] = 0
    dur[2, :] = 1
    lens = dur.sum(1)
    tgt_mask = torch.arange(int(lens.max())).unsqueeze(0) < lens.unsqueeze(1)
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([1, 2, 0]), "duration": dur, "tgt_mask": tgt_mask}
    r = tte_oracle.parrot_forward(sd, batch, dur_override=dur, tgt_mask=tgt_mask)
    dev_batch = _to_dev(batch)
    logits, _, mask_out, _ = m(dev_batch, inference=False)
    assert mask_out is dev_batch["tgt_mask"]
    assert torch.allclose(logits.cpu(), r["logits"], atol=3e-3, rtol=0), float((logits.cpu() - r["logits"]).abs().max())
    r2 = tte_oracle.parrot_forward(sd, batch, dur_override=dur)
    assert float((r2["logits"] - r["logits"]).abs().max()) > 1e-2          # the derived `<=` mask is a different computation


def test_parrot_rejects_overlong_like_reference():
    _, _, m = _parrot(seed=0, max_len=40)
    phones = torch.randint(2, VOCAB, (1, 40), generator=_gen(1))
    with pytest.raises(IndexError):                       # source length hits the pe table (fft.py:18)
        m({"phones": phones.to(DEV), "src_mask": (phones != 0).to(DEV), "speaker": torch.zeros(1, dtype=torch.long, device=DEV)}, inference=True)
    phones = torch.randint(2, VOCAB, (1, 30), generator=_gen(1))
    with pytest.raises(IndexError):                       # ~8 units per char -> unit length >= 40
        m({"phones": phones.to(DEV), "src_mask": (phones != 0).to(DEV), "speaker": torch.zeros(1, dtype=torch.long, device=DEV)}, inference=True)


def test_const8_duration_recipe():
    """Throughput recipe of SURVEY.md 8d: every real char -> exactly 8 units."""
    _, _, m = _parrot(seed=0, dur_mode="const8")
    phones = torch.randint(2, VOCAB, (2, 64), generator=_gen(4))
    phones[1, 40:] = 0
    ids, n = m.infer_device({"phones": phones.to(DEV), "src_mask": (phones != 0).to(DEV), "speaker": torch.tensor([0, 1], device=DEV)})
    assert ids.shape == (2, 512)
    assert n.tolist() == [512, 321]       # min(len + 1, T): the `<=` mask emits one extra position for the shorter item


def test_lit_parrot_checkpoint_roundtrip(tmp_path):
    cfg = _cfg()
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=3)
    lit = LitParrot(cfg, VOCAB, 0)
    lit.parrot.load_state_dict(sd)
    path = str(tmp_path / "epoch=0-step=1000.ckpt")
    lit.save_checkpoint(path)
    ck = torch.load(path, weights_only=True)
    assert set(ck) >= {"state_dict", "hyper_parameters"} and all(k.startswith("parrot.") for k in ck["state_dict"])
    model = LitParrot.load_from_checkpoint(path, weights_only=True).to(DEV)
    phones = torch.randint(2, VOCAB, (1, 12), generator=_gen(5))
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([3])}
    ids_ref, _ = tte_oracle.parrot_infer(sd, batch)
    assert model.infer(_to_dev(batch)) == ids_ref


def test_full_size_tensor_core_vs_fp32_path():
    """BASELINE configs[2] size (32 x 256 chars -> 2048 units): the tcgen05 path (gemm_mode 1) against the fp32 SIMT path
    (gemm_mode 0) of the same library - durations identical, unit ids identical except at genuine near-ties of the fp32
    logits (top-2 margin < 1e-4), and at most a handful of those."""
    cfg = _cfg()
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=0, dur_mode="const8")
    g = _gen(0)
    phones = torch.randint(2, VOCAB, (32, 256), generator=g).to(DEV)
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": (torch.arange(32) % NSPK).to(DEV)}
    out = {}
    for mode in (0, 1):
        m = Parrot(cfg, VOCAB, 0)
        m.load_state_dict(sd, strict=True)
        m = m.eval().to(DEV)
        m.gemm_mode = mode
        logits, _, tgt_mask, log_dur = m(batch, inference=True)
        out[mode] = (logits, tgt_mask, log_dur)
        del m
    assert out[0][0].shape == (32, 2048, 1000)
    assert torch.equal(out[0][1], out[1][1])                       # same durations -> same target mask
    assert torch.allclose(out[0][2], out[1][2], atol=1e-5, rtol=0)
    ids0, ids1 = out[0][0].argmax(-1), out[1][0].argmax(-1)
    bad = (ids0 != ids1).nonzero()
    top2 = torch.topk(out[0][0], 2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    assert bad.shape[0] <= 8, f"{bad.shape[0]} id mismatches of {ids0.numel()}"
    for b, t in bad.tolist():
        assert float(margin[b, t]) < 1e-4, f"id mismatch at a clear-cut position b={b} t={t} margin={float(margin[b, t])}"
    assert float((out[0][0] - out[1][0]).abs().max()) < 2e-3


def test_single_speaker_parrot_matches_oracle():
    """speakers.json with ONE entry: the reference builds no speaker_emb (modules/parrot.py:28-32); the handle is created with
    n_speaker = 0 and the speaker pointer is NULL.  Integer outputs bit-exact, as everywhere."""
    import json
    import tempfile
    d = tempfile.mkdtemp(prefix="parrot_one_spk_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({"only": 0}, f)
    cfg = configs.default_tte_config(root_path=d)
    sd = synth.synth_tte_state_dict(cfg, 150, 0, 1, seed=31)
    m = Parrot(cfg, 150, 0)
    assert m.speaker_emb is None
    m.load_state_dict(sd, strict=True)
    m = m.eval().to("cuda:0")
    g = torch.Generator().manual_seed(31)
    phones = torch.randint(2, 150, (3, 21), generator=g)
    phones[1, 9:] = 0
    batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.zeros(3, dtype=torch.long)}
    ids_ref, r = tte_oracle.parrot_infer(sd, batch)
    dev = {k: v.to("cuda:0") for k, v in batch.items()}
    logits, _, tgt_mask, log_dur = m(dev, inference=True)
    assert torch.equal(tgt_mask.cpu(), r["tgt_mask"])
    assert torch.allclose(log_dur.cpu(), r["log_dur"], atol=5e-5, rtol=0)
    assert torch.allclose(logits.cpu(), r["logits"], atol=3e-3, rtol=0)
    assert m.infer(dev) == ids_ref


@pytest.mark.parametrize("mode", [0, 1])
def test_infer_ragged_equals_every_item_run_alone(mode):
    """Parrot.infer_ragged: a padded batch in which every item is computed as the reference computes it ALONE (its inference mode,
    inference.py:31-37) - own pe rows (fft.py:18), zeros past its own end under the k = 9 / k = 3 convolutions, exactly its own unit
    count (no `<=` extra position).  Integer outputs bit-exact against the oracle run per item; logits BITWISE equal to this
    library's own batch-of-1 forward (same kernels, same accumulation order per row)."""
    cfg, sd, m = _parrot(seed=33)
    m.gemm_mode = mode
    g = _gen(12)
    lens = [37, 20, 5, 37, 1, 29]
    S = max(lens)
    phones = torch.zeros(len(lens), S, dtype=torch.long)
    for b, n in enumerate(lens):
        phones[b, :n] = torch.randint(2, VOCAB, (n,), generator=g)
    spk = torch.tensor([0, 9, 4, 2, 7, 1])
    ids, n_units, logits = m.infer_ragged(phones.to(DEV), torch.tensor(lens), spk.to(DEV), return_logits=True)
    n_units = n_units.tolist()
    for b, n in enumerate(lens):
        one = {"phones": phones[b:b + 1, :n], "src_mask": torch.ones(1, n, dtype=torch.bool), "speaker": spk[b:b + 1]}
        ids_ref, r = tte_oracle.parrot_infer(sd, one)
        assert n_units[b] == len(ids_ref[0]) == int(r["out_lens"][0])                  # durations bit-exact (integer path)
        assert ids[b, :n_units[b]].cpu().tolist() == ids_ref[0], b
        lg1, _, _, _ = m(_to_dev(one), inference=True)                                 # this library, batch of one
        assert torch.equal(logits[b, :n_units[b]], lg1[0]), (b, float((logits[b, :n_units[b]] - lg1[0]).abs().max()))
        assert torch.allclose(lg1[0].cpu(), r["logits"][0], atol=3e-3, rtol=0)
    # the reference's own padded-batch semantics (src_lens = NULL) still run on the same handle afterwards
    padded = m.infer({"phones": phones.to(DEV), "src_mask": (phones != 0).to(DEV), "speaker": spk.to(DEV)})
    ids_pad, _ = tte_oracle.parrot_infer(sd, {"phones": phones, "src_mask": phones != 0, "speaker": spk})
    assert [len(x) for x in padded] == [len(x) for x in ids_pad]


def test_fused_head_argmax_equals_argmax_of_the_logits():
    """Parrot.infer asks for ids only: the head GEMM takes the row argmax in its epilogue (64-bit atomicMax keys, first maximum on
    ties) and the logits never reach HBM.  Same accumulators, same bias add -> exactly torch.argmax of the logits Parrot.forward
    returns (reference modules/parrot.py:108,115), on a batch large enough for several N tiles, epilogue groups and CTA pairs."""
    cfg, sd, m = _parrot(seed=44, dur_mode="const8")
    g = _gen(13)
    phones = torch.randint(2, VOCAB, (5, 70), generator=g)
    phones[3, 41:] = 0
    batch = _to_dev({"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([0, 1, 2, 3, 4])})
    logits, _, tgt_mask, _ = m(batch, inference=True)
    ids, n_valid = m.infer_device(batch)
    assert ids.shape == logits.shape[:2] == (5, 560)
    assert torch.equal(ids, torch.argmax(logits, dim=-1))
    assert int(ids.max()) < 1000 and int(ids.min()) >= 0               # the 24 zero-padded head columns can never win
    # ties: with the head zeroed every logit is 0 -> the first index wins, as in torch.argmax
    with torch.no_grad():
        m.head.weight.zero_(); m.head.bias.zero_()
    ids0, _ = m.infer_device(batch)
    assert int(ids0.abs().max()) == 0
    with torch.no_grad():
        m.head.bias[7] = -0.0; m.head.bias[3] = 1.0; m.head.bias[900] = 1.0
    ids1, _ = m.infer_device(batch)
    assert torch.equal(ids1, torch.full_like(ids1, 3))


def test_infer_ragged_single_speaker_and_zero_duration_items():
    """Ragged mode on a single-speaker model (no speaker row to add: the encoder output is still zeroed past each item's end) and with
    an item whose characters all round to zero units."""
    import json
    import tempfile
    d = tempfile.mkdtemp(prefix="parrot_one_spk_ragged_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({"only": 0}, f)
    cfg = configs.default_tte_config(root_path=d)
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, 1, seed=35)
    m = Parrot(cfg, VOCAB, 0)
    m.load_state_dict(sd, strict=True)
    m = m.eval().to(DEV)
    g = _gen(14)
    lens = [11, 4, 17]
    phones = torch.zeros(3, 17, dtype=torch.long)
    for b, n in enumerate(lens):
        phones[b, :n] = torch.randint(2, VOCAB, (n,), generator=g)
    ids, n_units = m.infer_ragged(phones.to(DEV), torch.tensor(lens))
    for b, n in enumerate(lens):
        one = {"phones": phones[b:b + 1, :n], "src_mask": torch.ones(1, n, dtype=torch.bool), "speaker": torch.zeros(1, dtype=torch.long)}
        ids_ref, r = tte_oracle.parrot_infer(sd, one)
        assert int(n_units[b]) == len(ids_ref[0]) and ids[b, :int(n_units[b])].cpu().tolist() == ids_ref[0], b
    # durations forced to zero for one item: proj.bias very negative -> exp(.) - 1 rounds to 0 everywhere -> T = 0 for the whole batch
    with torch.no_grad():
        m.duration_predictor.proj.weight.zero_(); m.duration_predictor.proj.bias.fill_(-20.0)
    ids0, n0 = m.infer_ragged(phones.to(DEV), torch.tensor(lens))
    assert ids0.shape == (3, 0) and n0.tolist() == [0, 0, 0]


@pytest.mark.parametrize("mode,n_head", [(0, 2), (1, 2), (0, 4)])
def test_other_transformer_geometry_matches_reference_golden(golden_dir, mode, n_head):
    """A transformer shape the reference's default.yaml does not use (modules/parrot.py builds any of them from the config): 2 encoder /
    3 decoder layers, FFN 512 filters with a 5-tap first conv, 500 unit codes - against the vectors recorded from the reference's own
    module (oracle/gen_golden.py --other-geometries; 2 heads), and with 4 heads of 64 (fp32 path) against the oracle, which
    tests/test_oracle_vs_reference.py pins to the reference for that shape.  Ids identical wherever the reference's top-2 margin is
    clear, durations and masks identical."""
    g = np.load(os.path.join(golden_dir, "tte_geom_l23_s41.npz"))
    d = tempfile.mkdtemp(prefix="parrot_tte_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(NSPK)}, f)
    cfg = geometries.other_tte_config(n_head=n_head, root_path=d)
    sd = synth.synth_tte_state_dict(cfg, VOCAB, 0, NSPK, seed=int(g["seed"]))
    m = Parrot(cfg, VOCAB, 0)
    m.load_state_dict(sd, strict=True)
    m = m.eval().to(DEV)
    m.gemm_mode = mode
    batch = {"phones": torch.from_numpy(g["phones"]), "src_mask": torch.from_numpy(g["src_mask"]), "speaker": torch.from_numpy(g["speaker"])}
    ids = m.infer(_to_dev(batch))
    logits, _, tgt_mask, log_dur = m(_to_dev(batch), inference=True)
    assert logits.shape[-1] == 500
    if n_head == 2:      # the recorded reference run
        assert np.array_equal(tgt_mask.cpu().numpy(), g["tgt_mask"])
        np.testing.assert_allclose(log_dur.cpu().numpy(), g["log_dur"], rtol=0, atol=5e-5)
        stride = int(g["logits_row_stride"])
        np.testing.assert_allclose(logits[:, ::stride].cpu().numpy(), g["logits_rows"], rtol=0, atol=3e-3)
        assert [len(x) for x in ids] == g["ids_lens"].tolist()
        flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in ids])
        am = torch.argmax(logits, -1).cpu().numpy()
        bad = np.argwhere(am != g["argmax"])
        assert np.array_equal(flat, g["ids_flat"]) and bad.size == 0, [(tuple(ix), float(g["margin"][tuple(ix)])) for ix in bad[:5]]
        return
    ids_ref, r = tte_oracle.parrot_infer(sd, batch, n_head=n_head)
    assert tte_oracle._n_layers(sd, "encoder_layers") == 2 and tte_oracle._n_layers(sd, "decoder_layers") == 3
    assert torch.equal(tgt_mask.cpu(), r["tgt_mask"])
    assert torch.allclose(log_dur.cpu(), r["log_dur"], atol=5e-5, rtol=0)
    assert torch.allclose(logits.cpu(), r["logits"], atol=3e-3, rtol=0), float((logits.cpu() - r["logits"]).abs().max())
    top2 = torch.topk(r["logits"], 2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    for b, (a, e) in enumerate(zip(ids, ids_ref)):
        assert len(a) == len(e)
        for t, (u, v) in enumerate(zip(a, e)):
            assert u == v or margin[b, t] < 1e-4, (b, t, float(margin[b, t]))
