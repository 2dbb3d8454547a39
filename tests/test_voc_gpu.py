"""GPU parity: vocoder kernels and the full CodeGenerator / Generator against the CPU oracle and the
golden vectors recorded from the reference.  Tolerance on the waveform: 1e-3 abs (north_star).  Both
arithmetic modes are fp32-accurate - mode 0: fp32 SIMT kernels (observed ~3e-6), mode 1: tcgen05 implicit
GEMM on split-bf16 operands with fp32 accumulation (observed ~4e-5) - so the tests assert a tighter 2e-4.
Mode 2 (single-pass bf16, throughput only) is checked against a loose bound and is NOT a parity mode."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import geometries, voc_oracle
from parrot_tts_b200 import _lib, configs, ops, synth
from parrot_tts_b200.vocoder import CodeGenerator, Generator

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
WAV_TOL = 1e-3      # contract
WAV_TIGHT = 2e-4    # what fp32 kernels should actually achieve


def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("K", [3, 7, 11])
@pytest.mark.parametrize("D", [1, 3, 5])
@pytest.mark.parametrize("C,L", [(16, 1531), (32, 700), (64, 333), (128, 130)])
def test_conv1d_kernel(K, D, C, L):
    g = _gen(K * 100 + D * 10 + C)
    B = 2
    x = torch.randn(B, C, L, generator=g)
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g)
    r = torch.randn(B, C, L, generator=g)
    ref = F.conv1d(F.leaky_relu(x, 0.1), w, b, padding=(K - 1) * D // 2, dilation=D)
    got = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), dilation=D, pre_slope=0.1).cpu()
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())
    ref2 = F.leaky_relu(F.conv1d(x, w, b, padding=(K - 1) * D // 2, dilation=D) + r, 0.1)
    got2 = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), dilation=D, res=r.to(DEV), post_slope=0.1).cpu()
    assert torch.allclose(got2, ref2, atol=2e-5, rtol=1e-5), float((got2 - ref2).abs().max())


def test_conv1d_pre_shape():
    # conv_pre: 256 -> 512, k7
    g = _gen(5)
    x = torch.randn(3, 256, 45, generator=g)
    w = torch.randn(512, 256, 7, generator=g) / (256 * 7) ** 0.5
    b = torch.randn(512, generator=g)
    ref = F.conv1d(x, w, b, padding=3)
    got = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=3e-5, rtol=1e-5), float((got - ref).abs().max())


@pytest.mark.parametrize("K,U", [(11, 5), (8, 4), (4, 2)])
@pytest.mark.parametrize("Cin,Cout,L", [(512, 256, 37), (64, 32, 300), (32, 16, 1030)])
def test_conv_transpose1d_kernel(K, U, Cin, Cout, L):
    g = _gen(K * 10 + U + Cin)
    x = torch.randn(2, Cin, L, generator=g)
    w = torch.randn(Cin, Cout, K, generator=g) / (Cin * K / U) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=U, padding=(K - U) // 2)
    got = ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), stride=U, pre_slope=0.1).cpu()
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())


MODES = [0, 1]


@pytest.mark.parametrize("C,K,D,L", [(16, 3, 1, 300), (16, 11, 5, 5000), (32, 7, 3, 777), (64, 11, 1, 500), (64, 1, 1, 129),
                                     (128, 3, 5, 260), (256, 7, 3, 200), (256, 11, 5, 130),
                                     # CTA pairs with the weights split and resident: 64 channels at k = 7 / 11, 128 channels at k = 3
                                     (64, 7, 5, 1000), (64, 11, 5, 700), (64, 11, 3, 257), (128, 3, 1, 600), (64, 3, 3, 900)])
def test_conv1d_tensor_core_kernel(C, K, D, L):
    g = _gen(K * 100 + D * 10 + C + 1)
    B = 2
    x = torch.randn(B, C, L, generator=g)
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g)
    r = torch.randn(B, C, L, generator=g)
    ref = F.leaky_relu(F.conv1d(F.leaky_relu(x, 0.1), w, b, padding=(K - 1) * D // 2, dilation=D) + r, 0.1)
    got = ops.conv1d_tc(x.to(DEV), w.to(DEV), b.to(DEV), dilation=D, res=r.to(DEV), pre_slope=0.1, post_slope=0.1).cpu()
    # split-bf16 operands: ~2^-16 relative per product; values are O(5)
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), float((got - ref).abs().max())
    fast = ops.conv1d_tc(x.to(DEV), w.to(DEV), b.to(DEV), dilation=D, res=r.to(DEV), pre_slope=0.1, post_slope=0.1, single_pass=True).cpu()
    assert torch.allclose(fast, ref, atol=8e-2, rtol=2e-2), float((fast - ref).abs().max())


def test_conv1d_tensor_core_wide_n():
    # conv_pre shape: 256 -> 512 (two N tiles of 256), k7
    g = _gen(77)
    x = torch.randn(2, 256, 301, generator=g)
    w = torch.randn(512, 256, 7, generator=g) / (256 * 7) ** 0.5
    b = torch.randn(512, generator=g)
    ref = F.conv1d(x, w, b, padding=3)
    got = ops.conv1d_tc(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), float((got - ref).abs().max())


@pytest.mark.parametrize("Cin,Cout,K,U,L", [(512, 256, 11, 5, 150), (256, 128, 8, 4, 129), (128, 64, 8, 4, 200), (64, 32, 4, 2, 300), (32, 16, 4, 2, 1031)])
def test_conv_transpose1d_tensor_core_kernel(Cin, Cout, K, U, L):
    g = _gen(Cin + K)
    x = torch.randn(2, Cin, L, generator=g)
    w = torch.randn(Cin, Cout, K, generator=g) / (Cin * K / U) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=U, padding=(K - U) // 2)
    got = ops.conv_transpose1d_tc(x.to(DEV), w.to(DEV), b.to(DEV), stride=U, pre_slope=0.1).cpu()
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), float((got - ref).abs().max())


def _code_generator(seed, weight_norm=True, mode=1):
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=seed, weight_norm=True)
    gen = CodeGenerator(h)
    gen.conv_mode = mode
    gen.load_state_dict(sd, strict=True)
    if not weight_norm:
        gen.remove_weight_norm()
    return h, sd, gen.to(DEV)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["voc_b1_t24", "voc_b2_t40", "voc_cfg1_t512"])
def test_code_generator_matches_reference_golden(golden_dir, name, mode):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    _, _, gen = _code_generator(int(g["seed"]), mode=mode)
    wav = gen(code=torch.from_numpy(g["code"]).to(DEV), spkr=torch.from_numpy(g["spkr"]).to(DEV)).cpu().numpy()
    assert wav.shape == g["wav"].shape
    err = np.abs(wav - g["wav"]).max()
    assert err < WAV_TOL, err
    assert err < WAV_TIGHT, err


@pytest.mark.parametrize("mode", MODES)
def test_generator_matches_reference_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "gen_b1_t12.npz"))
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=True)
    sd = {k: v for k, v in sd.items() if not k.startswith(("dict.", "spkr."))}
    gen = Generator(h)
    gen.conv_mode = mode
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(DEV)
    wav = gen(torch.from_numpy(g["x"]).to(DEV)).cpu().numpy()
    assert np.abs(wav - g["wav"]).max() < WAV_TIGHT


@pytest.mark.parametrize("mode", MODES)
def test_remove_weight_norm_same_audio(mode):
    h, sd, gen = _code_generator(7, mode=mode)
    code = torch.randint(0, 1000, (2, 19), generator=_gen(1))
    spkr = torch.tensor([[1], [4]])
    a = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    gen.remove_weight_norm()
    assert "conv_pre.weight" in gen.state_dict() and "conv_pre.weight_g" not in gen.state_dict()
    b = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    assert (a - ref).abs().max() < WAV_TIGHT
    assert (b - ref).abs().max() < WAV_TIGHT


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("B,T", [(1, 1), (1, 3), (3, 50), (2, 131)])
def test_code_generator_vs_oracle_rectangle(B, T, mode):
    """lens=None: padded-rectangle semantics, identical to the reference run on the same batch."""
    h, sd, gen = _code_generator(11, mode=mode)
    g = _gen(B * 1000 + T)
    code = torch.randint(0, 1000, (B, T), generator=g)
    spkr = torch.randint(0, 10, (B, 1), generator=g)
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    got = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    assert got.shape == (B, 1, 320 * T)
    assert (got - ref).abs().max() < WAV_TIGHT, float((got - ref).abs().max())


@pytest.mark.parametrize("mode", MODES)
def test_code_generator_ragged_equals_batch_of_one(mode):
    """lens given: each item equals the reference's own batch-of-1 result (its inference mode,
    utils/vocoder/inference.py:149), and is exactly zero past its end."""
    h, sd, gen = _code_generator(13, mode=mode)
    g = _gen(99)
    lens = [57, 23, 40, 1]
    T = max(lens)
    code = torch.randint(0, 1000, (len(lens), T), generator=g)
    spkr = torch.randint(0, 10, (len(lens), 1), generator=g)
    got = gen(code=code.to(DEV), spkr=spkr.to(DEV), lens=torch.tensor(lens)).cpu()
    for i, n in enumerate(lens):
        ref = voc_oracle.code_generator_forward(sd, h, code[i:i + 1, :n], spkr[i:i + 1])
        assert (got[i:i + 1, :, :320 * n] - ref).abs().max() < WAV_TIGHT
        assert got[i, :, 320 * n:].abs().max() == 0 if n < T else True


def test_bf16_throughput_mode_is_close_but_not_parity():
    h, sd, gen = _code_generator(11, mode=2)
    g = _gen(4)
    code = torch.randint(0, 1000, (2, 40), generator=g)
    spkr = torch.randint(0, 10, (2, 1), generator=g)
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    got = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    assert (got - ref).abs().max() < 5e-2      # plain bf16 products: ~1e-2 at full scale, outside the 1e-3 contract


def test_full_size_properties_tensor_core():
    """BASELINE configs[1] size (8 x 1024 units): properties that do not need the CPU oracle at full size -
    the batch is item-wise independent (item b alone == item b in the batch) and a 1024-unit item equals
    its 256-unit chunks away from the seams (the generator is convolutional, receptive radius < 32 units)."""
    h, sd, gen = _code_generator(0, mode=1)
    g = _gen(8)
    code = torch.randint(0, 1000, (8, 1024), generator=g).to(DEV)
    spkr = torch.randint(0, 10, (8, 1), generator=g).to(DEV)
    full = gen(code=code, spkr=spkr)
    assert full.shape == (8, 1, 327680) and bool(torch.isfinite(full).all())
    one = gen(code=code[3:4], spkr=spkr[3:4])
    assert torch.equal(one[0], full[3])                      # same tiles, same arithmetic: bit-identical
    halo = 32
    chunk = gen(code=code[5:6, 256 - halo:512 + halo], spkr=spkr[5:6])
    a = chunk[0, 0, halo * 320:(halo + 256) * 320]
    b = full[5, 0, 256 * 320:512 * 320]
    assert (a - b).abs().max() < 2e-4


def test_pcm16_matches_reference_cast():
    y = (torch.rand(1, 1, 4096, generator=_gen(3)) * 2 - 1)
    y[0, 0, :4] = torch.tensor([0.99999, -1.0, 1.5 / 32768, -1.5 / 32768])
    ref = voc_oracle.to_int16(y)
    got = ops.wav_to_pcm16(y.to(DEV)).cpu().squeeze().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape", [(1,), (7,), (4099,), (3, 40961), (2, 163840)])
def test_peak_normalize_matches_librosa_restatement(shape):
    """librosa.util.normalize of the int16 audio (utils/vocoder/inference.py:169): bit-exact float32 result per utterance."""
    g = np.random.default_rng(sum(shape))
    pcm = g.integers(-9000, 9000, size=shape, dtype=np.int16)
    if len(shape) == 2:
        pcm[1] = 0                                            # an all-zero utterance is returned unchanged
        pcm[0, 5] = -32768                                    # |-32768| does not fit int16: the peak is taken in int32
    ref = np.stack([voc_oracle.peak_normalize(r) for r in pcm]) if len(shape) == 2 else voc_oracle.peak_normalize(pcm)
    got = ops.peak_normalize(torch.from_numpy(pcm).to(DEV)).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, ref)
    if len(shape) == 1:
        from parrot_tts_b200.vocoder import normalize
        assert np.array_equal(normalize(pcm), ref)            # numpy in, numpy out: the reference's call shape


def test_no_cpu_fallback():
    h, sd, gen = _code_generator(0)
    with pytest.raises(Exception):
        gen.cpu()(code=torch.zeros(1, 4, dtype=torch.long), spkr=torch.zeros(1, 1, dtype=torch.long))


def test_streamed_chunks_equal_unchunked():
    """BASELINE configs[4]: chunked synthesis with a 32-unit halo reproduces the unchunked waveform (the generator's
    receptive-field radius is ~23 units)."""
    from parrot_tts_b200.vocoder import stream_chunks, stream_generate
    h = configs.default_vocoder_config()
    gen = CodeGenerator(h)
    gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=3))
    gen = gen.to(DEV)
    g = torch.Generator().manual_seed(11)
    T = 700
    code = torch.randint(0, 1000, (1, T), generator=g).to(DEV)
    spkr = torch.tensor([[4]]).to(DEV)
    full = gen(code=code, spkr=spkr)[:, 0]
    parts = list(stream_generate(gen, code, spkr, chunk_units=256, halo_units=32))
    assert [p[0] for p in parts] == [a * 320 for _, _, a, _ in stream_chunks(T, 256, 32)]
    got = torch.cat([w for _, w in parts], dim=1)
    assert got.shape == full.shape
    err = float((got - full).abs().max())
    assert err <= 1e-6, err
    # a halo shorter than the receptive field must NOT reproduce it (the test would be vacuous otherwise)
    bad = torch.cat([w for _, w in stream_generate(gen, code, spkr, chunk_units=256, halo_units=2)], dim=1)
    assert float((bad - full).abs().max()) > 1e-5


def test_single_cta_path_matches_torch():
    """The single-CTA variant of the wide tc_conv layers (PB_TC_CLUSTER=0, read once per process; the default is the
    cta_group::2 pair path) in a subprocess: wide conv with streamed weights (C = 128, k = 7, dilation 3, residual) and a wide
    GEMM (TTE-sized), against torch fp32."""
    import subprocess
    import sys
    code = r'''
import torch, torch.nn.functional as F
from parrot_tts_b200 import ops
g = torch.Generator().manual_seed(5)
C, L, K, D = 128, 700, 7, 3
x = torch.randn(2, C, L, generator=g); w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
b = torch.randn(C, generator=g); r = torch.randn(2, C, L, generator=g)
ref = F.conv1d(x, w, b, dilation=D, padding=(K - 1) * D // 2) + r
got = ops.conv1d_tc(x.cuda(), w.cuda(), b.cuda(), dilation=D, res=r.cuda()).cpu()
e1 = float((got - ref).abs().max())
xm = torch.randn(300, 256, generator=g); wm = torch.randn(768, 256, generator=g) / 16
e2 = float((ops.linear(xm.cuda(), wm.cuda(), None, mode=1).cpu() - xm @ wm.t()).abs().max())
print("ERR", e1, e2)
assert e1 < 1e-4 and e2 < 2e-5, (e1, e2)
'''
    env = dict(os.environ, PB_TC_CLUSTER="0", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


def test_batch_synthesizer_equals_per_utterance_generate():
    """Batching front door: ragged length-bucketed batches, device-side PCM and peak normalisation, long-form
    streaming - every utterance equals the reference-style one-at-a-time `generate()` result (bit-exact int16)."""
    from parrot_tts_b200.serving import BatchSynthesizer
    from parrot_tts_b200.vocoder import generate, normalize
    h, sd, gen = _code_generator(0)
    g = _gen(11)
    lengths = [40, 37, 36, 12, 0, 150, 33, 9]
    codes = [torch.randint(0, 1000, (n,), generator=g) for n in lengths]
    spk = [int(torch.randint(0, 10, (1,), generator=g)) for _ in lengths]
    synth = BatchSynthesizer(gen, max_batch=3, max_pad_ratio=1.25, long_form_units=100, chunk_units=64, halo_units=32)
    got = synth(codes, spk)
    assert synth.last_plan["long_form"] == [5] and synth.last_plan["empty"] == [4]
    assert all(len(b) <= 3 for b in synth.last_plan["batches"])
    got_n = synth(codes, spk, normalize=True)
    for i, n in enumerate(lengths):
        assert got[i].dtype == np.int16 and got[i].shape == (n * 320,)
        if n == 0:
            continue
        ref, _ = generate(h, gen, dict(code=codes[i].view(1, -1).to(DEV), spkr=torch.tensor([[spk[i]]], device=DEV)))
        if i == 5:      # streamed in chunks: the unchunked values up to fp32 round-off (waveform within 2e-4 -> a few LSB at most)
            assert np.abs(got[i].astype(np.int32) - ref.astype(np.int32)).max() <= 2
        else:
            assert np.array_equal(got[i], ref), i
        assert np.array_equal(got_n[i], normalize(got[i]))
    part = synth(codes, spk, indices=[0, 3])
    assert [p is not None for p in part] == [True, False, False, True, False, False, False, False]
    assert np.array_equal(part[3], got[3])


def test_cuda_graph_replay_equals_direct_forward():
    """GraphedGenerator: the captured forward (PDL edges, cluster launches, TMA descriptors as kernel parameters) replays
    bit-identically for new inputs, ragged lengths included; streaming through graphs equals streaming without."""
    from parrot_tts_b200.vocoder import GraphedGenerator, stream_generate
    h, sd, gen = _code_generator(0)
    g = _gen(21)
    gg = GraphedGenerator(gen, 2, 48)
    for rep in range(3):
        code = torch.randint(0, 1000, (2, 48), generator=g).to(DEV)
        spkr = torch.randint(0, 10, (2, 1), generator=g).to(DEV)
        lens = None if rep == 0 else torch.tensor([48, 17 + rep])
        want = gen(code=code, spkr=spkr) if lens is None else gen(code=code, spkr=spkr, lens=lens)
        got = gg(code, spkr, lens).clone()
        assert torch.equal(got, want), rep
    with pytest.raises(ValueError):
        gg(torch.zeros(2, 47, dtype=torch.long, device=DEV), spkr)
    code = torch.randint(0, 1000, (1, 300), generator=g).to(DEV)
    spkr = torch.randint(0, 10, (1, 1), generator=g).to(DEV)
    plain = torch.cat([w for _, w in stream_generate(gen, code, spkr, chunk_units=128, halo_units=32)], dim=1)
    graphs = {}
    viag = torch.cat([w.clone() for _, w in stream_generate(gen, code, spkr, chunk_units=128, halo_units=32, graphs=graphs)], dim=1)
    assert sorted(graphs) == [(76, False), (160, False), (192, False)] and torch.equal(plain, viag)
    ahead = list(stream_generate(gen, code, spkr, chunk_units=128, halo_units=32, lookahead=1))
    assert [o for o, _ in ahead] == [0, 128 * 320, 256 * 320] and torch.equal(torch.cat([w for _, w in ahead], dim=1), plain)
    with pytest.raises(ValueError):
        list(stream_generate(gen, code, spkr, chunk_units=128, halo_units=32, graphs=graphs, lookahead=1))
    on_host = list(stream_generate(gen, code, spkr, chunk_units=128, halo_units=32, lookahead=1, host=True, pcm16=True))
    want = ops.wav_to_pcm16(plain).cpu().numpy()
    assert all(isinstance(w, np.ndarray) and w.dtype == np.int16 for _, w in on_host)
    assert np.array_equal(np.concatenate([w for _, w in on_host], axis=1), want)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_pcm16_epilogue_equals_cast_of_the_waveform(mode):
    """generate()'s tail (x 32768, int16 truncation; reference utils/vocoder/inference.py:71-73) as conv_post's epilogue: bitwise the
    cast of the very waveform the plain forward returns - rectangular and ragged batches, and through generate() itself."""
    from parrot_tts_b200.vocoder import generate
    h, sd, gen = _code_generator(0, mode=mode)
    g = _gen(31)
    for B, T, lens in ((1, 9, None), (3, 41, None), (3, 57, [57, 20, 1])):
        code = torch.randint(0, 1000, (B, T), generator=g).to(DEV)
        spkr = torch.randint(0, 10, (B, 1), generator=g).to(DEV)
        kw = dict(code=code, spkr=spkr)
        if lens is not None:
            kw["lens"] = torch.tensor(lens, dtype=torch.int32)
        wav = gen(**kw)
        pcm = gen.forward_pcm16(**kw)
        assert pcm.dtype == torch.int16 and pcm.shape == wav.shape
        assert np.array_equal(pcm.cpu().numpy(), voc_oracle.to_int16(wav.cpu()).reshape(pcm.shape)), (B, T)
        assert torch.equal(pcm, ops.wav_to_pcm16(wav))
    code = torch.randint(0, 1000, (1, 33), generator=g).to(DEV)
    spkr = torch.randint(0, 10, (1, 1), generator=g).to(DEV)
    audio, _ = generate(h, gen, dict(code=code, spkr=spkr))
    assert np.array_equal(audio, voc_oracle.to_int16(gen(code=code, spkr=spkr).cpu()).squeeze())


def test_out_of_range_ids_raise_like_nn_embedding():
    """nn.Embedding raises IndexError for an id outside its table (reference utils/vocoder/models.py:155-156); padded positions
    (past an item's length) are never looked up."""
    h, sd, gen = _code_generator(1, mode=1)
    code = torch.randint(0, 1000, (2, 12), generator=_gen(5)).to(DEV)
    spkr = torch.tensor([[1], [2]], device=DEV)
    gen(code=code, spkr=spkr)
    gen.check_ids()
    bad = code.clone()
    bad[1, 7] = 1000                                   # the TTE dataset's pad value for codes (= num_embeddings)
    gen(code=bad, spkr=spkr, lens=torch.tensor([12, 7], dtype=torch.int32))
    gen.check_ids()                                    # position 7 of item 1 is padding under lens
    gen(code=bad, spkr=spkr)
    with pytest.raises(IndexError):
        gen.check_ids()
    gen.check_ids()                                    # the flag is cleared by the report
    gen(code=code, spkr=torch.tensor([[1], [10]], device=DEV))
    with pytest.raises(IndexError):
        gen.check_ids()


def test_cuda_graph_is_recaptured_when_its_pointers_die():
    """A graph records raw pointers into the engine's arena and workspace: after a larger forward regrew the workspace, or after the
    weights changed (new engine), the next call captures again instead of replaying into freed memory."""
    from parrot_tts_b200.vocoder import GraphedGenerator
    h, sd, gen = _code_generator(1, mode=1)
    g = _gen(41)
    gg = GraphedGenerator(gen, 1, 16)
    code = torch.randint(0, 1000, (1, 16), generator=g).to(DEV)
    spkr = torch.randint(0, 10, (1, 1), generator=g).to(DEV)
    want = gen(code=code, spkr=spkr)
    assert torch.equal(gg(code, spkr), want) and gg.captures == 1
    big = torch.randint(0, 1000, (4, 400), generator=g).to(DEV)
    gen(code=big, spkr=torch.zeros(4, 1, dtype=torch.long, device=DEV))      # regrows the workspace: old pointers are dead
    assert torch.equal(gg(code, spkr), want) and gg.captures == 2
    assert torch.equal(gg(code, spkr), want) and gg.captures == 2            # stable afterwards
    torch.nn.Linear(3, 3)                                                      # unrelated registration: nothing of this model changed
    assert torch.equal(gg(code, spkr), want) and gg.captures == 2
    with torch.no_grad():
        gen.conv_post.bias.add_(0.25)                                          # weights changed: new engine, new arena
    want2 = gen(code=code, spkr=spkr)
    assert not torch.equal(want2, want)
    assert torch.equal(gg(code, spkr), want2) and gg.captures == 3


def test_speech_pipeline_equals_one_utterance_at_a_time():
    """serving.SpeechPipeline (char tokens -> TTE per utterance -> device hand-over of the ids -> ragged vocoder batches -> int16):
    every utterance equals Parrot.infer + generate() run on it alone; a rank share (indices=) returns only its utterances."""
    import json
    import tempfile
    from oracle import tte_oracle
    from parrot_tts_b200.serving import SpeechPipeline, rank_share_tte
    from parrot_tts_b200.tte import Parrot
    from parrot_tts_b200.vocoder import generate
    h, sd, gen = _code_generator(1, mode=1)
    d = tempfile.mkdtemp(prefix="parrot_pipe_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    cfg = configs.default_tte_config(root_path=d)
    sd_t = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=12)
    tte = Parrot(cfg, 150, 0)
    tte.load_state_dict(sd_t)
    tte = tte.eval().to(DEV)
    g = _gen(51)
    phones = [torch.randint(2, 150, (n,), generator=g) for n in (5, 9, 3, 14, 7)]
    spk = [1, 4, 0, 9, 2]
    pipe = SpeechPipeline(tte, gen, max_batch=2)
    audio = pipe(phones, spk)
    for i, ph in enumerate(phones):
        batch = {"phones": ph.view(1, -1), "src_mask": ph.view(1, -1) != 0, "speaker": torch.tensor([spk[i]])}
        ids_ref, _ = tte_oracle.parrot_infer(sd_t, batch)
        assert pipe.last_units[i].cpu().tolist() == ids_ref[0]                      # integer path: bit-exact vs the oracle
        ref, _ = generate(h, gen, dict(code=torch.tensor([ids_ref[0]], device=DEV), spkr=torch.tensor([[spk[i]]], device=DEV)))
        assert np.array_equal(audio[i], ref), i
    assert sorted(i for b in pipe.last_batches for i in b) == list(range(5)) and all(len(b) <= 2 for b in pipe.last_batches)
    f32 = pipe(phones, spk, pcm16=False, to_host=False)                              # raw waveform, left on the device
    for i in range(5):
        assert f32[i].is_cuda and f32[i].dtype == torch.float32
        assert np.array_equal(voc_oracle.to_int16(f32[i].cpu()), audio[i])
    nrm = pipe(phones, spk, normalize=True)
    from parrot_tts_b200.vocoder import normalize as _normalize
    assert np.array_equal(nrm[3], _normalize(audio[3]))
    lens = [int(u.numel()) for u in pipe.last_units]
    shares = [rank_share_tte(lens, 2, r) for r in range(2)]
    assert sorted(shares[0] + shares[1]) == list(range(5))
    part = pipe(phones, spk, indices=shares[1])
    assert [p is not None for p in part] == [i in shares[1] for i in range(5)]
    for i in shares[1]:
        assert np.array_equal(part[i], audio[i])


def test_mel_spectrogram_matches_reference_golden(golden_dir):
    """Device log-mel (pb_mel_spectrogram) against the output of the reference's own function (utils/vocoder/dataset.py:43-69)
    recorded in tests/golden; tolerance 2e-3 in the log domain (fp32 radix-2 FFT vs torch.stft; the basis itself is parity-unpinned)."""
    from parrot_tts_b200.vocoder.mel import mel_spectrogram
    g = np.load(os.path.join(golden_dir, "mel_b2_n16384.npz"))
    got = mel_spectrogram(torch.from_numpy(g["y"]).to(DEV), 1024, 80, 16000, 256, 1024, 0, 8000).cpu().numpy()
    assert got.shape == g["mel"].shape
    err = np.abs(got - g["mel"])
    assert err.max() < 2e-3, float(err.max())
    with pytest.raises(NotImplementedError):
        mel_spectrogram(torch.zeros(1, 4096, device=DEV), 1024, 80, 16000, 256, 1024, 0, 8000, center=True)


def test_mel_distance_of_generated_audio_vs_oracle():
    """The metric in use: log-mel of the waveform this library generates vs log-mel of the oracle's waveform for the same units -
    the mel-L1 the reference's training loop reports (train.py:135-137), here as a parity figure beside the sample tolerance."""
    from oracle import mel_oracle
    from parrot_tts_b200.vocoder.mel import mel_spectrogram
    h, sd, gen = _code_generator(0, mode=1)
    g = _gen(61)
    code = torch.randint(0, 1000, (2, 40), generator=g)
    spkr = torch.randint(0, 10, (2, 1), generator=g)
    wav = gen(code=code.to(DEV), spkr=spkr.to(DEV))
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    m_dev = mel_spectrogram(wav.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax).cpu()
    m_ref = mel_oracle.mel_spectrogram(ref.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax)
    assert m_dev.shape == m_ref.shape
    assert float((m_dev - m_ref).abs().mean()) < 1e-3, float((m_dev - m_ref).abs().mean())


def test_other_hifigan_geometry_matches_reference_golden(golden_dir, mode=1):
    """A HiFi-GAN shape the reference config does not use (utils/vocoder/models.py builds any of them from `h`): three upsampling
    stages x16, two ResBlock kernels 5 / 9 with dilations (1, 2, 4) / (1, 3, 5) - against the vector recorded from the reference's own
    module (oracle/gen_golden.py --other-geometries), and item by item against the oracle in ragged mode.  Exercises the kernels at
    other tap counts and dilations than 3 / 7 / 11 x (1, 3, 5): the 64-channel layers as single CTAs (k = 5) and CTA pairs (k = 9), the
    fused 32-channel pairs, the 16-channel pairs tap-by-tap (k = 5) and time-grouped (k = 9), a mean over two ResBlocks.  The
    tensor-core path (the default) takes any odd kernel size; the fp32 SIMT validation path is compiled for the reference's
    3 / 7 / 11 x (1, 3, 5) only and says so."""
    g = np.load(os.path.join(golden_dir, "voc_geom_x16_t300.npz"))
    h = geometries.other_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=True)
    gen = CodeGenerator(h)
    gen.conv_mode = mode
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(DEV)
    code, spkr = torch.from_numpy(g["code"]), torch.from_numpy(g["spkr"])
    lens = [300, 41, 1]
    got = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    assert got.shape == g["wav"].shape == (3, 1, 16 * max(lens))
    err = float((got - torch.from_numpy(g["wav"])).abs().max())
    assert err < WAV_TIGHT, err
    rag = gen(code=code.to(DEV), spkr=spkr.to(DEV), lens=torch.tensor(lens)).cpu()
    for i, n in enumerate(lens):
        alone = voc_oracle.code_generator_forward(sd, h, code[i:i + 1, :n], spkr[i:i + 1])
        assert (rag[i:i + 1, :, :16 * n] - alone).abs().max() < WAV_TIGHT
        assert n == max(lens) or rag[i, :, 16 * n:].abs().max() == 0
    simt = CodeGenerator(h)
    simt.conv_mode = 0
    simt.load_state_dict(sd, strict=True)
    with pytest.raises(_lib.ParrotB200Error, match="no sm_100a instantiation"):
        simt.to(DEV)(code=code.to(DEV), spkr=spkr.to(DEV))
