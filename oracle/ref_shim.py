"""ORACLE support (test infrastructure): import the *reference's own* modules from /root/reference.

Works only in the build container (the GPU box has no /root/reference); used by
oracle/gen_golden.py and tests/test_oracle_vs_reference.py to pin the restatements in
tte_oracle.py / voc_oracle.py to the real thing.  Nothing is copied: the reference files are
imported from where they lie, with the two shims described in SURVEY.md section 8c:
  * ``speakers.json`` in a temp ``root_path`` (reference modules/parrot.py:24-26 reads it);
  * a stub ``matplotlib`` and ``utils/vocoder`` first on sys.path so that
    ``from utils import init_weights, get_padding, AttrDict`` (utils/vocoder/models.py:8) resolves.
"""
import importlib
import json
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("PARROT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "modules", "parrot.py"))


def _purge(names):
    for n in list(sys.modules):
        if n in names or any(n.startswith(p + ".") for p in names):
            del sys.modules[n]


def load_reference_parrot(cfg, src_vocab_size, src_pad_idx, n_speaker):
    """Instantiate the reference Parrot (eval mode)."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        _purge({"modules", "utils"})
        sys.path.insert(0, REF_ROOT)
        parrot_mod = importlib.import_module("modules.parrot")
        d = tempfile.mkdtemp(prefix="parrot_ref_")
        with open(os.path.join(d, "speakers.json"), "w") as f:
            json.dump({f"spk{i}": i for i in range(n_speaker)}, f)
        import copy
        cfg = copy.deepcopy(cfg)
        cfg["path"]["root_path"] = d
        model = parrot_mod.Parrot(cfg, src_vocab_size, src_pad_idx).eval()
        return model
    finally:
        sys.path[:] = saved_path
        _purge({"modules", "utils"})
        for k, v in saved_mods.items():
            sys.modules.setdefault(k, v)


def load_reference_vocoder(h):
    """Instantiate the reference CodeGenerator (weight norm still attached, as in demo cell 13)."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        _purge({"modules", "utils", "models"})
        if "matplotlib" not in sys.modules:
            mpl = types.ModuleType("matplotlib")
            mpl.use = lambda *a, **k: None
            pylab = types.ModuleType("matplotlib.pylab")
            mpl.pylab = pylab
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pylab"] = pylab
        sys.path.insert(0, os.path.join(REF_ROOT, "utils", "vocoder"))
        models = importlib.import_module("models")
        ref_utils = importlib.import_module("utils")
        gen = models.CodeGenerator(ref_utils.AttrDict(dict(h)))
        return gen
    finally:
        sys.path[:] = saved_path
        _purge({"modules", "utils", "models"})
        for k, v in saved_mods.items():
            sys.modules.setdefault(k, v)


def load_reference_mel_spectrogram(mel_fn):
    """The reference's own ``mel_spectrogram`` (utils/vocoder/dataset.py:43-69), imported from where it lies.  Its module imports
    amfm_decompy, soundfile and librosa at the top (none installed, none used by this function except ``librosa.filters.mel``):
    they are stubbed, and ``librosa.filters.mel`` is ``mel_fn`` - so everything but the mel BASIS (padding, STFT, magnitude, log)
    is the reference's code; the basis is the oracle's restatement of the published librosa algorithm (parity-unpinned)."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        _purge({"dataset", "librosa", "amfm_decompy", "soundfile"})
        for name in ("amfm_decompy", "amfm_decompy.basic_tools", "amfm_decompy.pYAAPT", "soundfile", "librosa", "librosa.filters", "librosa.util"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["amfm_decompy"].basic_tools = sys.modules["amfm_decompy.basic_tools"]
        sys.modules["amfm_decompy"].pYAAPT = sys.modules["amfm_decompy.pYAAPT"]
        sys.modules["librosa"].filters = sys.modules["librosa.filters"]
        sys.modules["librosa"].util = sys.modules["librosa.util"]
        sys.modules["librosa.filters"].mel = mel_fn
        sys.modules["librosa.util"].normalize = lambda x, **k: x
        sys.path.insert(0, os.path.join(REF_ROOT, "utils", "vocoder"))
        ds = importlib.import_module("dataset")
        return ds.mel_spectrogram
    finally:
        sys.path[:] = saved_path
        _purge({"dataset", "librosa", "amfm_decompy", "soundfile"})
        for k, v in saved_mods.items():
            sys.modules.setdefault(k, v)
