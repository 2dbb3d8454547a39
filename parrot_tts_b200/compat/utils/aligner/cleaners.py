"""reference utils/aligner/cleaners.py"""
from parrot_tts_b200.text.cleaners import *  # noqa: F401,F403
from parrot_tts_b200.text.cleaners import (english_cleaners, nonenglish_cleaners,  # noqa: F401
                                           nonenglish_cleaners_no_transliteration)
