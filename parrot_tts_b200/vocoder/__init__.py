from .inference import MAX_WAV_VALUE, GraphedGenerator, generate, load_checkpoint, normalize, stream_chunks, stream_generate  # noqa: F401
from .models import CodeGenerator, Generator, ResBlock1, ResBlock2  # noqa: F401
from .utils import AttrDict, get_padding, init_weights  # noqa: F401
from .mel import mel_filterbank, mel_spectrogram  # noqa: F401
