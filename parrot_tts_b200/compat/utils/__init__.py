"""Package mode (``utils.vocoder.models``) AND script mode (``from utils import init_weights, get_padding, AttrDict``, reference
utils/vocoder/models.py:8) both resolve - in the reference tree only one of the two does."""
from parrot_tts_b200.vocoder.utils import AttrDict, get_padding, init_weights  # noqa: F401
