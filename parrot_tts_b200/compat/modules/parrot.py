"""reference modules/parrot.py"""
from parrot_tts_b200.tte.parrot import Parrot  # noqa: F401
