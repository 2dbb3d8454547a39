"""Import-path shims: the reference's module paths, resolving to this package.

    import parrot_tts_b200.compat as compat; compat.install()

puts this directory first on ``sys.path`` so that the import lines of demo.ipynb cell 0 and of the reference scripts -
``from modules.data import DFATokenizer``, ``from inference import LitParrot``, ``from utils.vocoder.models import CodeGenerator``,
``from utils.aligner.cleaners import english_cleaners``, and the script-mode ``from utils import AttrDict`` /
``from models import CodeGenerator`` of utils/vocoder/inference.py:24-27 - run unchanged and bind the sm_100a backend.
Each shim module is a re-export, nothing more.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_TOP_LEVEL = ("modules", "utils", "inference", "models", "dataset")


def install():
    """Idempotent.  Modules of the same names imported earlier (the reference's own, say) are dropped from ``sys.modules``."""
    for name in list(sys.modules):
        if name.split(".")[0] in _TOP_LEVEL and not getattr(sys.modules[name], "__file__", "").startswith(HERE):
            del sys.modules[name]
    if HERE in sys.path:
        sys.path.remove(HERE)
    sys.path.insert(0, HERE)
    return HERE
