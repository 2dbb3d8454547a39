"""Hyper-parameter sources of the two stages, kept in the reference's own formats.

TTE:     nested dict as loaded from ``utils/TTE/TTE_config.yaml`` (reference
         utils/TTE/TTE_config.yaml:1-44; consumed at modules/parrot.py:15-18,36-65).
Vocoder: ``AttrDict`` as loaded from ``utils/vocoder/config.json`` (reference
         utils/vocoder/config.json:1-44; consumed at utils/vocoder/models.py:73-91,125-130).

Only the keys the inference hot path reads are materialised here; the yaml/json
files shipped next to this module carry the full set for scripts that parse them.
"""
import copy
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
TTE_CONFIG_YAML = os.path.join(_HERE, "configs", "TTE_config.yaml")
VOCODER_CONFIG_JSON = os.path.join(_HERE, "configs", "config.json")


class AttrDict(dict):
    """dict whose keys are also attributes (reference utils/vocoder/utils.py:77-80)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


_TTE_DEFAULT = {
    "path": {"root_path": "runs/TTE", "alignment_path": "runs/aligner"},
    "preprocess": {"val_size": 100, "hubert_codes": 1000, "speaker": "_"},
    "transformer": {
        "encoder": {"n_layer": 4, "n_head": 2, "dropout_p": 0.1},
        "decoder": {"n_layer": 4, "n_head": 2, "dropout_p": 0.1},
        "d_model": 256,
        "conv_n_filter": 1024,
        "conv_kernel_sizes": [9, 1],
        "max_len": 3500,
    },
    "duration_predictor": {"n_filter": 256, "kernel_size": 3, "dropout_p": 0.5},
}

_VOC_DEFAULT = {
    "resblock": "1",
    "upsample_rates": [5, 4, 4, 2, 2],
    "upsample_kernel_sizes": [11, 8, 8, 4, 4],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "num_embeddings": 1000,
    "embedding_dim": 128,
    "model_in_dim": 256,
    "code_hop_size": 320,
    "multispkr": "_",
    "sampling_rate": 16000,
    # mel_spectrogram parameters (reference utils/vocoder/config.json:26-36; validation metric, vocoder/mel.py)
    "num_mels": 80, "n_fft": 1024, "hop_size": 256, "win_size": 1024, "fmin": 0, "fmax": 8000,
}


def default_tte_config(root_path=None, max_len=None):
    cfg = copy.deepcopy(_TTE_DEFAULT)
    if root_path is not None:
        cfg["path"]["root_path"] = root_path
    if max_len is not None:
        cfg["transformer"]["max_len"] = int(max_len)
    return cfg


def default_vocoder_config():
    return AttrDict(copy.deepcopy(_VOC_DEFAULT))


def load_vocoder_config(path=VOCODER_CONFIG_JSON):
    with open(path) as f:
        return AttrDict(json.loads(f.read()))


def load_tte_config(path=TTE_CONFIG_YAML):
    import yaml
    with open(path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)
