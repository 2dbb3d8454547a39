"""reference utils/vocoder/utils.py:32-45,77-80"""
from parrot_tts_b200.vocoder.utils import AttrDict, get_padding, init_weights  # noqa: F401
