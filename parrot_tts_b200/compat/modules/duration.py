"""reference modules/duration.py"""
from parrot_tts_b200.tte.duration import Conv, DurationPredictor, length_regulator  # noqa: F401
