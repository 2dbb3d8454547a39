"""Helpers the reference's vocoder modules import from ``utils`` (reference utils/vocoder/utils.py:32-45,77-80)."""
from ..configs import AttrDict  # noqa: F401


def get_padding(kernel_size, dilation=1):
    """'same' padding of a dilated kernel: (k*d - d) / 2 (reference utils/vocoder/utils.py:44-45)."""
    return (kernel_size * dilation - dilation) // 2


def init_weights(m, mean=0.0, std=0.01):
    """N(mean, std) init of every Conv* module's weight (reference utils/vocoder/utils.py:32-35)."""
    if "Conv" in type(m).__name__:
        m.weight.data.normal_(mean, std)
