"""Host-side helpers of the TTE path (reference modules/data.py:8-61): masks and the DFA tokenizer.
Pure host logic - no device math lives here."""
import pickle
from pathlib import Path

import torch


def get_mask_from_lengths(lengths, max_len=None, device=None):
    """bool [B, max_len], True where ``position <= length`` - i.e. length+1 positions are marked valid.
    The off-by-one is the reference's (modules/data.py:18) and is part of the parity contract: the
    extra zero-row position is a live attention key and ``Parrot.infer`` emits its unit id."""
    lengths = torch.as_tensor(lengths)
    if max_len is None:
        max_len = int(lengths.max())
    pos = torch.arange(int(max_len))
    if device:
        lengths, pos = lengths.to(device), pos.to(device)
    return pos.unsqueeze(0) <= lengths.unsqueeze(1)


def get_mask_from_batch(batch, pad_idx):
    return batch != pad_idx


class DFATokenizer:
    """Symbol table produced by the aligner (``symbols.pkl``), with ``<pad>``/``<sep>`` prepended and the
    blank character mapped to ``sil`` (reference modules/data.py:28-61)."""
    pad = "<pad>"
    sep = "<sep>"

    def __init__(self, alignment_path):
        with open(Path(alignment_path) / "symbols.pkl", "rb") as f:
            table = pickle.load(f)
        if isinstance(table, dict):
            symbols = list(table.keys())
        elif isinstance(table, list):
            symbols = list(table)
        else:
            raise TypeError(f"symbols.pkl holds a {type(table).__name__}, expected dict or list")
        if " " in symbols:                        # the blank is the aligner's silence token (first occurrence, as in the reference)
            symbols[symbols.index(" ")] = "sil"
        self.symbols = [self.pad, self.sep] + symbols
        self.stoi = {s: i for i, s in enumerate(self.symbols)}
        self.itos = dict(enumerate(self.symbols))
        self.pad_idx = self.stoi[self.pad]
        self.sep_idx = self.stoi[self.sep]

    def __len__(self):
        return len(self.symbols)

    def tokenize(self, phoneme_seq):
        return [self.stoi[s] for s in phoneme_seq]
