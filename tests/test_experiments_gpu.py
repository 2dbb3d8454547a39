"""Parity of the PREPARED (round-2) switches against the default path - bitwise, since none of them changes the arithmetic.

Skipped unless PB_TEST_EXPERIMENTS=1: the switches were written after round 1's GPU budget was spent and have not run on a GPU yet
(tools/gpu_exp_r2.sh).  Each case runs in a fresh interpreter because the library reads its switches once.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PB_TEST_EXPERIMENTS") != "1", reason="prepared experiments: set PB_TEST_EXPERIMENTS=1")]

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
    return _run(tmp_path_factory.mktemp("exp"), "default", {"PB_VOC_CHAINS": None, "PB_TC_TMA_OUT": None, "PB_LIB_VARIANT": None})


def test_chain_streams_are_bit_identical(tmp_path, baseline):
    got = _run(tmp_path, "chains", {"PB_VOC_CHAINS": "1"})
    _same(baseline, got)
    assert np.array_equal(got["graph"], got["graph_again"])          # fork / join survives CUDA-graph capture and replay


def test_tma_store_on_k3_pair_layers_is_bit_identical(tmp_path, baseline):
    _same(baseline, _run(tmp_path, "tma3", {"PB_TC_TMA_OUT": "3"}))


VARIANTS = {   # name -> (build flags, environment); build each with `python -m parrot_tts_b200.build --variant NAME --define ...`
    "episplit": ("--define PB_TC_EPI_SPLIT --define PB_POST_TILE4", [{"PB_TC_EPI_SPLIT": "1"}, {"PB_TC_EPI_SPLIT": "2"}]),
    "family": ("--define PB_TC_FAMILY", [{}]),
    "r2all": ("--define PB_TC_EPI_SPLIT --define PB_POST_TILE4 --define PB_TC_FAMILY", [{"PB_TC_EPI_SPLIT": "1", "PB_VOC_CHAINS": "1"}]),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variant_builds_are_bit_identical(tmp_path, baseline, variant):
    flags, envs = VARIANTS[variant]
    lib = os.path.join(ROOT, "parrot_tts_b200", "lib", f"libparrot_b200_{variant}.so")
    if not os.path.isfile(lib):
        pytest.skip(f"build it first: python -m parrot_tts_b200.build --variant {variant} {flags}")
    for n, env in enumerate(envs):
        _same(baseline, _run(tmp_path, f"{variant}{n}", dict(env, PB_LIB_VARIANT=variant)))
