"""GPU parity: vocoder kernels and the full CodeGenerator / Generator against the CPU oracle and the
golden vectors recorded from the reference.  Tolerance on the waveform: 1e-3 abs (north_star); the
kernels are fp32 so the observed error is ~1e-5 and the tests assert a tighter 2e-4."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import voc_oracle
from parrot_tts_b200 import configs, ops, synth
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


def _code_generator(seed, weight_norm=True):
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=seed, weight_norm=True)
    gen = CodeGenerator(h)
    gen.load_state_dict(sd, strict=True)
    if not weight_norm:
        gen.remove_weight_norm()
    return h, sd, gen.to(DEV)


@pytest.mark.parametrize("name", ["voc_b1_t24", "voc_b2_t40"])
def test_code_generator_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    _, _, gen = _code_generator(int(g["seed"]))
    wav = gen(code=torch.from_numpy(g["code"]).to(DEV), spkr=torch.from_numpy(g["spkr"]).to(DEV)).cpu().numpy()
    assert wav.shape == g["wav"].shape
    err = np.abs(wav - g["wav"]).max()
    assert err < WAV_TOL, err
    assert err < WAV_TIGHT, err


def test_generator_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gen_b1_t12.npz"))
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=True)
    sd = {k: v for k, v in sd.items() if not k.startswith(("dict.", "spkr."))}
    gen = Generator(h)
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(DEV)
    wav = gen(torch.from_numpy(g["x"]).to(DEV)).cpu().numpy()
    assert np.abs(wav - g["wav"]).max() < WAV_TIGHT


def test_remove_weight_norm_same_audio():
    h, sd, gen = _code_generator(7)
    code = torch.randint(0, 1000, (2, 19), generator=_gen(1))
    spkr = torch.tensor([[1], [4]])
    a = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    gen.remove_weight_norm()
    assert "conv_pre.weight" in gen.state_dict() and "conv_pre.weight_g" not in gen.state_dict()
    b = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    assert (a - ref).abs().max() < WAV_TIGHT
    assert (b - ref).abs().max() < WAV_TIGHT


@pytest.mark.parametrize("B,T", [(1, 1), (1, 3), (3, 50), (2, 131)])
def test_code_generator_vs_oracle_rectangle(B, T):
    """lens=None: padded-rectangle semantics, identical to the reference run on the same batch."""
    h, sd, gen = _code_generator(11)
    g = _gen(B * 1000 + T)
    code = torch.randint(0, 1000, (B, T), generator=g)
    spkr = torch.randint(0, 10, (B, 1), generator=g)
    ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
    got = gen(code=code.to(DEV), spkr=spkr.to(DEV)).cpu()
    assert got.shape == (B, 1, 320 * T)
    assert (got - ref).abs().max() < WAV_TIGHT, float((got - ref).abs().max())


def test_code_generator_ragged_equals_batch_of_one():
    """lens given: each item equals the reference's own batch-of-1 result (its inference mode,
    utils/vocoder/inference.py:149), and is exactly zero past its end."""
    h, sd, gen = _code_generator(13)
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


def test_pcm16_matches_reference_cast():
    y = (torch.rand(1, 1, 4096, generator=_gen(3)) * 2 - 1)
    y[0, 0, :4] = torch.tensor([0.99999, -1.0, 1.5 / 32768, -1.5 / 32768])
    ref = voc_oracle.to_int16(y)
    got = ops.wav_to_pcm16(y.to(DEV)).cpu().squeeze().numpy()
    assert np.array_equal(got, ref)


def test_no_cpu_fallback():
    h, sd, gen = _code_generator(0)
    with pytest.raises(Exception):
        gen.cpu()(code=torch.zeros(1, 4, dtype=torch.long), spkr=torch.zeros(1, 1, dtype=torch.long))
