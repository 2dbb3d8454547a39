"""reference modules/data.py:8-61"""
from parrot_tts_b200.tte.data import DFATokenizer, get_mask_from_batch, get_mask_from_lengths  # noqa: F401
