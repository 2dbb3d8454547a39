from .data import DFATokenizer, get_mask_from_batch, get_mask_from_lengths  # noqa: F401
from .duration import DurationPredictor, length_regulator  # noqa: F401
from .fft import Attention, ConvLayer, FFTBlock, SinusoidalPosEmb  # noqa: F401
from .parrot import LitParrot, Parrot  # noqa: F401
