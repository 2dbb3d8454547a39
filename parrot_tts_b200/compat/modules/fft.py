"""reference modules/fft.py"""
from parrot_tts_b200.tte.fft import Attention, ConvLayer, FFTBlock, SinusoidalPosEmb  # noqa: F401
