"""ORACLE support (test infrastructure): two model geometries the reference's config files do not use.  The reference builds any of them
from the same config objects (modules/parrot.py:22-53, utils/vocoder/models.py:72-93); oracle/gen_golden.py records its outputs for
them and the CPU / GPU tests hold the oracle and the CUDA path to those vectors."""
from parrot_tts_b200 import configs


def other_tte_config(n_head=2, root_path=None):
    """2 encoder / 3 decoder layers, FFN 512 filters x 5 taps, 500 unit codes (default.yaml: 4 / 4, 1024 x 9, 1000)."""
    cfg = configs.default_tte_config(root_path=root_path)
    tr = cfg["transformer"]
    tr["encoder"]["n_layer"], tr["decoder"]["n_layer"] = 2, 3
    tr["encoder"]["n_head"] = tr["decoder"]["n_head"] = n_head
    tr["conv_n_filter"], tr["conv_kernel_sizes"] = 512, [5, 1]
    cfg["preprocess"]["hubert_codes"] = 500
    return cfg


def other_vocoder_config():
    """x16 in three stages, ResBlock kernels 5 / 9 with dilations (1, 2, 4) / (1, 3, 5) (config.json: x320 in five, 3 / 7 / 11 x (1, 3, 5))."""
    h = configs.default_vocoder_config()
    h.update({"upsample_rates": [4, 2, 2], "upsample_kernel_sizes": [8, 4, 4], "upsample_initial_channel": 128,
              "resblock_kernel_sizes": [5, 9], "resblock_dilation_sizes": [[1, 2, 4], [1, 3, 5]], "code_hop_size": 16})
    return h
