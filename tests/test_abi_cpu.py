"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/parrot_b200.h declares, the ctypes table mirrors the header one to one, and the product path
refuses to run without CUDA (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest
import torch

from parrot_tts_b200 import _lib, build, configs, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "parrot_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_declares_symbols():
    names = _declared()
    assert "pb_voc_forward" in names and "pb_tte_encode" in names and len(names) >= 30


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header(lib_path):
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.pb_version() == 3
    assert lib.pb_last_error() is not None


def test_struct_layouts_match_header():
    # pb_voc_cfg: 3 + 8 + 8 + 1 + 4 + 1 + 16 + 3 int32 ; pb_tte_cfg: 16 int32
    assert ctypes.sizeof(_lib.VocCfg) == 4 * (3 + 8 + 8 + 1 + 4 + 1 + 16 + 4)
    assert ctypes.sizeof(_lib.TteCfg) == 4 * 16


def test_error_paths_without_gpu(lib_path):
    lib = _lib.load()
    assert lib.pb_voc_forward(None, None, None, None, None, 1, 1, None, None) == _lib.PB_ERR_INVALID
    assert b"null" in lib.pb_last_error()
    assert lib.pb_tte_decode(None, 1, 1, 1, None, None, None, None, None) == _lib.PB_ERR_INVALID
    with pytest.raises(IndexError):
        _lib.check(_lib.PB_ERR_INDEX)
    with pytest.raises(_lib.ParrotB200Error):
        _lib.check(_lib.PB_ERR_CUDA)


def test_product_has_no_cpu_path():
    from parrot_tts_b200.vocoder import CodeGenerator
    h = configs.default_vocoder_config()
    gen = CodeGenerator(h)
    gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0), strict=True)
    with pytest.raises(_lib.ParrotB200Error):
        gen(code=torch.zeros(1, 4, dtype=torch.long), spkr=torch.zeros(1, 1, dtype=torch.long))


def test_product_never_imports_the_oracle():
    """Only tests/, oracle/ itself, bench.py (CPU legs) and __graft_entry__.smoke() may use the oracle - not the package, not tools/."""
    for top in ("parrot_tts_b200", "tools"):
        for dp, _, fs in os.walk(os.path.join(ROOT, top)):
            for f in fs:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_state_dict_layouts_match_reference_names():
    """Key names/shapes of SURVEY.md section 8b (old-style weight norm for the vocoder; TTE keys)."""
    from parrot_tts_b200.vocoder import CodeGenerator
    h = configs.default_vocoder_config()
    sd = CodeGenerator(h).state_dict()
    assert tuple(sd["conv_pre.weight_v"].shape) == (512, 256, 7) and tuple(sd["conv_pre.weight_g"].shape) == (512, 1, 1)
    assert tuple(sd["ups.0.weight_g"].shape) == (512, 1, 1) and tuple(sd["ups.0.weight_v"].shape) == (512, 256, 11)
    assert tuple(sd["resblocks.14.convs2.2.weight_v"].shape) == (16, 16, 11)
    assert tuple(sd["conv_post.weight_v"].shape) == (1, 16, 7)
    assert tuple(sd["dict.weight"].shape) == (1000, 128) and tuple(sd["spkr.weight"].shape) == (10, 128)
    assert set(sd) == set(synth.synth_vocoder_state_dict(h, seed=0))


def test_checkpoint_files_load_like_the_reference(tmp_path):
    """The two checkpoint files of the demo: a Lightning .ckpt dict (extra trainer keys ignored, `parrot.`
    prefix) through LitParrot.load_from_checkpoint(path, weights_only=True), and torch.load(path)['generator'] for the vocoder."""
    import json
    import torch
    from parrot_tts_b200.tte import LitParrot
    from parrot_tts_b200.vocoder import CodeGenerator, load_checkpoint
    with open(tmp_path / "speakers.json", "w") as f:
        json.dump({f"s{i}": i for i in range(3)}, f)
    cfg = configs.default_tte_config(root_path=str(tmp_path))
    sd = synth.synth_tte_state_dict(cfg, 40, 0, 3, seed=2)
    ckpt = {"epoch": 7, "global_step": 1234, "pytorch-lightning_version": "2.4.0", "optimizer_states": [{"state": {}}], "lr_schedulers": [],
            "hyper_parameters": {"data_config": cfg, "src_vocab_size": 40, "src_pad_idx": 0},
            "state_dict": {"parrot." + k: v for k, v in sd.items()}}
    torch.save(ckpt, tmp_path / "tte.ckpt")
    model = LitParrot.load_from_checkpoint(str(tmp_path / "tte.ckpt"), weights_only=True)
    assert not model.training and model.parrot.src_vocab_size == 40
    got = model.parrot.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(Exception):                       # no silent CPU execution of the loaded model
        model.infer({"phones": torch.ones(1, 4, dtype=torch.long), "src_mask": torch.ones(1, 4, dtype=torch.bool),
                     "speaker": torch.zeros(1, dtype=torch.long)})

    h = configs.default_vocoder_config()
    sd_v = synth.synth_vocoder_state_dict(h, seed=2)
    torch.save({"generator": sd_v}, tmp_path / "g_00000001")
    gen = CodeGenerator(h)
    gen.load_state_dict(load_checkpoint(str(tmp_path / "g_00000001"))["generator"])
    got = gen.state_dict()
    assert all(torch.equal(got[k], sd_v[k]) for k in sd_v)


def test_reference_import_paths_resolve_to_this_package():
    """demo.ipynb cell 0 / utils/vocoder/inference.py:24-27 import lines, unchanged, after compat.install() (fresh interpreter)."""
    import subprocess
    import sys
    code = """
import parrot_tts_b200.compat as compat
compat.install(); compat.install()
from modules.data import DFATokenizer, get_mask_from_batch
from utils.aligner.cleaners import english_cleaners, nonenglish_cleaners, nonenglish_cleaners_no_transliteration
from inference import LitParrot
from utils.vocoder.utils import AttrDict
from utils.vocoder.models import CodeGenerator
from utils.vocoder.dataset import MAX_WAV_VALUE
from modules.parrot import Parrot
from modules.fft import FFTBlock, SinusoidalPosEmb
from modules.duration import DurationPredictor, length_regulator
from modules import Parrot as P2
from utils import init_weights, get_padding, AttrDict as A2      # script mode of utils/vocoder/models.py:8
from models import CodeGenerator as C2                            # script mode of utils/vocoder/inference.py:27
from dataset import MAX_WAV_VALUE as M2
from utils.vocoder.inference import generate
import parrot_tts_b200.tte, parrot_tts_b200.vocoder, parrot_tts_b200.text
assert CodeGenerator is C2 is parrot_tts_b200.vocoder.CodeGenerator and LitParrot is parrot_tts_b200.tte.LitParrot
assert Parrot is P2 is parrot_tts_b200.tte.Parrot and AttrDict is A2 and MAX_WAV_VALUE == M2 == 32768.0
assert english_cleaners is parrot_tts_b200.text.english_cleaners and english_cleaners('Dr. No 2') == 'doctor no two'
assert get_padding(7, 3) == 9 and AttrDict(a=1).a == 1
print('ok')
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_header_is_plain_c_and_links_from_c(tmp_path, lib_path):
    """The boundary is a C ABI: include/parrot_b200.h compiles as C99 (pedantic), and a C program linked against the library gets
    the version, a thread-local error message and error codes - no GPU, no Python, no C++ in between."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "parrot_b200.h"
int main(void) {
    pb_voc* v = NULL;
    pb_voc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    if (pb_version() != PB_ABI_VERSION) return 1;
    if (pb_voc_forward(NULL, NULL, NULL, NULL, NULL, 1, 1, NULL, NULL) >= 0) return 2;      /* null handle */
    if (strlen(pb_last_error()) == 0) return 3;
    if (pb_voc_create(&cfg, &v) >= 0) return 4;                                              /* an all-zero config is rejected */
    if (pb_wav_to_pcm16(NULL, 4, NULL, NULL) >= 0) return 5;
    printf("%d %s\n", pb_version(), pb_last_error());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(lib_path)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lparrot_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("3 ")
