"""Bitwise parity of the library's dev switches against the default path - none of them changes the arithmetic:
  PB_VOC_CHAINS=0/1   the three ResBlock chains of a stage on three streams (default: on for forwards of <= 2048 units)
  PB_TC_TMA_OUT=2     TMA-store epilogue on every eligible layer (default 1: the one-CTA-per-SM layers)
  PB_TC_EPI_SPLIT=0/2 column-split epilogue groups off / on every one-CTA-per-SM layer (default 1: the CTA-pair layers)
  PB_TC_CLUSTER=0     single-CTA path instead of CTA pairs
  PB_TC_PAIR64=0      the 64-channel k = 7 / 11 layers as one CTA per SM instead of CTA pairs
  PB_TC_CL_RESIDENT=0 CTA pairs always stream their weights (no resident halves)
  PB_TP_GROUPS=3/4    tiles in flight in the fused ResBlock-step kernel of the 32 / 16-channel stages (default: per layer)
  PB_TP_TOEPLITZ=0/1  time-grouped conv2 in the 16-channel fused kernel never / for every kernel size (default: k >= 7)
  PB_TP_RES_SMEM=0    the tap-by-tap fused kernel re-reads the residual rows from global memory instead of the x tile in shared memory
  PB_TP_TMAE=0        ... with its residual / output rows as row-per-lane global accesses instead of shared-memory staging + TMA
Each case runs in a fresh interpreter because the library reads its switches once.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
SWITCHES = ("PB_VOC_CHAINS", "PB_TC_TMA_OUT", "PB_TC_EPI_SPLIT", "PB_TC_CLUSTER", "PB_TC_PAIR64", "PB_TC_CL_RESIDENT", "PB_TP_GROUPS", "PB_TP_TOEPLITZ", "PB_TP_TMAE", "PB_TP_RES_SMEM", "PB_LIB_VARIANT")

CODE = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator, GraphedGenerator
h = configs.default_vocoder_config()
gen = CodeGenerator(h)
gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0))
gen = gen.to("cuda:0")
g = torch.Generator().manual_seed(11)
out = {{}}
for B, T in ((1, 7), (2, 50), (8, 300)):
    code = torch.randint(0, 1000, (B, T), generator=g).cuda()
    spkr = torch.randint(0, 10, (B, 1), generator=g).cuda()
    out[f"rect_{{B}}_{{T}}"] = gen(code=code, spkr=spkr).cpu().numpy()
code = torch.randint(0, 1000, (3, 120), generator=g).cuda()
spkr = torch.randint(0, 10, (3, 1), generator=g).cuda()
lens = torch.tensor([120, 33, 1], dtype=torch.int32).cuda()
out["ragged"] = gen(code=code, spkr=spkr, lens=lens).cpu().numpy()
gg = GraphedGenerator(gen, 3, 120)
out["graph"] = gg(code, spkr).cpu().numpy().copy()
out["graph_again"] = gg(code, spkr).cpu().numpy().copy()
np.savez({path!r}, **out)
"""


def _run(tmp_path, name, env_extra):
    path = str(tmp_path / f"{name}.npz")
    env = {k: v for k, v in os.environ.items() if k not in env_extra}          # None = make sure the variable is unset
    env.update({k: v for k, v in env_extra.items() if v is not None})
    r = subprocess.run([sys.executable, "-c", CODE.format(root=ROOT, path=path)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return dict(np.load(path))


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


@pytest.fixture(scope="module")
def baseline(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("exp"), "default", {k: None for k in SWITCHES})


def test_chain_streams_are_bit_identical(tmp_path, baseline):
    # the default runs the small cases with chain streams and the 8 x 300-unit one without: forced off / on must both agree with it
    _same(baseline, _run(tmp_path, "nochains", {"PB_VOC_CHAINS": "0"}))
    got = _run(tmp_path, "chains", {"PB_VOC_CHAINS": "1"})
    _same(baseline, got)
    assert np.array_equal(got["graph"], got["graph_again"])          # fork / join survives CUDA-graph capture and replay


def test_tma_store_epilogue_everywhere_is_bit_identical(tmp_path, baseline):
    _same(baseline, _run(tmp_path, "tma2", {"PB_TC_TMA_OUT": "2"}))


@pytest.mark.parametrize("mode", ["0", "2"])
def test_column_split_epilogue_modes_are_bit_identical(tmp_path, baseline, mode):
    _same(baseline, _run(tmp_path, "episplit" + mode, {"PB_TC_EPI_SPLIT": mode}))


@pytest.mark.parametrize("switch", ["PB_TC_PAIR64", "PB_TC_CL_RESIDENT", "PB_TC_CLUSTER"])
def test_cta_pair_placement_switches_are_bit_identical(tmp_path, baseline, switch):
    # where a layer runs (one CTA per SM, a CTA pair streaming its weights, a pair with resident halves) changes neither the order of
    # the K loop nor the products: same bits
    _same(baseline, _run(tmp_path, switch.lower(), {switch: "0"}))


@pytest.mark.parametrize("groups", ["3", "4"])
def test_fused_pair_groups_are_bit_identical(tmp_path, baseline, groups):
    # how many tiles the fused conv1 -> conv2 kernel keeps in flight is scheduling only
    _same(baseline, _run(tmp_path, "tpgroups" + groups, {"PB_TP_GROUPS": groups}))


@pytest.mark.parametrize("mode", ["0", "1"])
def test_time_grouped_conv2_is_bit_identical(tmp_path, baseline, mode):
    # the block-Toeplitz operand only adds exact zeros to each output's sum and keeps the order of its taps: same bits as the
    # tap-by-tap kernel, at every kernel size
    _same(baseline, _run(tmp_path, "toeplitz" + mode, {"PB_TP_TOEPLITZ": mode}))


def test_time_grouped_kernel_epilogue_paths_are_bit_identical(tmp_path, baseline):
    # residual rows in / output rows out through shared memory and TMA, or straight from / to global memory: the same arithmetic
    _same(baseline, _run(tmp_path, "tmae0", {"PB_TP_TMAE": "0", "PB_TP_TOEPLITZ": "1"}))
    _same(baseline, _run(tmp_path, "tmae1", {"PB_TP_TMAE": "1", "PB_TP_TOEPLITZ": "1"}))


def test_residual_from_the_x_tile_is_bit_identical(tmp_path, baseline):
    # the x tile a fused pair loaded for conv1 holds the residual rows of its outputs: same values as the global re-read
    _same(baseline, _run(tmp_path, "ressmem0", {"PB_TP_RES_SMEM": "0"}))
    _same(baseline, _run(tmp_path, "ressmem0t0", {"PB_TP_RES_SMEM": "0", "PB_TP_TOEPLITZ": "0"}))
    _same(baseline, _run(tmp_path, "ressmem2t0", {"PB_TP_RES_SMEM": "2", "PB_TP_TOEPLITZ": "0"}))
