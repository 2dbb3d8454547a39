"""Host text front end (reference utils/aligner/cleaners.py, demo.ipynb cells 9-11): dependency-free cleaners and the
text -> batch helpers.  String work only."""
from .cleaners import (basic_cleaners, english_cleaners, nonenglish_cleaners,  # noqa: F401
                       nonenglish_cleaners_no_transliteration)
from .frontend import TextToSpeech, characters_to_batch, looks_english, pick_cleaner, text_to_characters  # noqa: F401
