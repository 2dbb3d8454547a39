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
    "groups4": ("--define PB_TC_GROUPS4 --define PB_TC_EPI_SPLIT", [{"PB_TC_EPI_SPLIT": "1"}, {"PB_TC_EPI_SPLIT": "0"}]),
    "r2all": ("--define PB_TC_EPI_SPLIT --define PB_POST_TILE4 --define PB_TC_FAMILY --define PB_TC_GROUPS4", [{"PB_TC_EPI_SPLIT": "1", "PB_VOC_CHAINS": "1"}]),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variant_builds_are_bit_identical(tmp_path, baseline, variant):
    flags, envs = VARIANTS[variant]
    lib = os.path.join(ROOT, "parrot_tts_b200", "lib", f"libparrot_b200_{variant}.so")
    if not os.path.isfile(lib):
        pytest.skip(f"build it first: python -m parrot_tts_b200.build --variant {variant} {flags}")
    for n, env in enumerate(envs):
        _same(baseline, _run(tmp_path, f"{variant}{n}", dict(env, PB_LIB_VARIANT=variant)))


# ---- coverage written without GPU access (same gate; to be moved into the regular GPU files once they have run) -------------------
def test_single_speaker_parrot_matches_oracle():
    """speakers.json with ONE entry: the reference builds no speaker_emb (modules/parrot.py:28-32); the handle is created with
    n_speaker = 0 and the speaker pointer is NULL.  Integer outputs bit-exact, as everywhere."""
    import json
    import tempfile
    import torch
    from oracle import tte_oracle
    from parrot_tts_b200 import configs, synth
    from parrot_tts_b200.tte import Parrot
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


def test_text_to_speech_end_to_end(tmp_path):
    """text -> cleaners -> DFATokenizer -> Parrot.infer (per utterance) -> BatchSynthesizer -> int16 PCM, against the same chain on
    the oracle: unit ids bit-exact, PCM within the 1e-3 waveform bar (32.768 int16 steps, +1 for the truncation)."""
    import json
    import pickle
    import torch
    from oracle import tte_oracle, voc_oracle
    from parrot_tts_b200 import configs, synth
    from parrot_tts_b200.serving import BatchSynthesizer
    from parrot_tts_b200.text import TextToSpeech
    from parrot_tts_b200.tte import DFATokenizer, Parrot
    from parrot_tts_b200.vocoder import CodeGenerator
    symbols = list("abcdefghijklmnopqrstuvwxyz ,.")
    with open(tmp_path / "symbols.pkl", "wb") as f:
        pickle.dump(symbols, f)
    with open(tmp_path / "speakers.json", "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    tok = DFATokenizer(tmp_path)
    cfg = configs.default_tte_config(root_path=str(tmp_path))
    sd_t = synth.synth_tte_state_dict(cfg, len(tok), tok.pad_idx, 10, seed=4)
    tte = Parrot(cfg, len(tok), tok.pad_idx)
    tte.load_state_dict(sd_t)
    tte = tte.eval().to("cuda:0")
    h = configs.default_vocoder_config()
    sd_v = synth.synth_vocoder_state_dict(h, seed=4)
    gen = CodeGenerator(h)
    gen.load_state_dict(sd_v)
    gen = gen.to("cuda:0")
    tts = TextToSpeech(tok, symbols, tte.infer, BatchSynthesizer(gen), device="cuda:0")
    texts, speakers = ["Dr. Who has 2 cats.", "???", "hello there, general"], [1, 2, 3]
    audio = tts(texts, speakers)
    assert audio[1] is None and tts.last_units[1] is None
    for i in (0, 2):
        phones = torch.tensor([tok.tokenize(tts.last_characters[i])])
        ids_ref, _ = tte_oracle.parrot_infer(sd_t, {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([speakers[i]])})
        assert tts.last_units[i] == ids_ref[0]
        ref = voc_oracle.code_generator_forward(sd_v, h, torch.tensor([ids_ref[0]]), torch.tensor([[speakers[i]]]))
        pcm_ref = voc_oracle.to_int16(ref)
        assert audio[i].dtype == np.int16 and audio[i].shape == pcm_ref.shape
        assert int(np.abs(audio[i].astype(np.int32) - pcm_ref.astype(np.int32)).max()) <= 34
