"""Duration predictor and length regulator with the reference's names and parameter layout
(reference modules/duration.py), executing on libparrot_b200.so."""
import torch
import torch.nn as nn

from .. import _lib
from ._engine import TteEngine, need_cuda


def length_regulator(batch_seq, batch_dur, tgt_mask=None):
    """Repeat every row of ``batch_seq`` [B,S,D] ``batch_dur`` [B,S] times, zero-pad to the longest item
    (reference duration.py:6-24).  Returns (expanded [B,T,D], tgt_mask bool [B,T]) with the reference's
    ``position <= length`` mask when none is given."""
    need_cuda(batch_seq, "length_regulator")
    dev = batch_seq.device
    B, S, D = batch_seq.shape
    seq = batch_seq.to(torch.float32).contiguous()
    dur = batch_dur.to(device=dev, dtype=torch.int64).contiguous()
    T = int(dur.sum(dim=1).max())          # data-dependent output length: one device->host sync, as in the reference
    if tgt_mask is not None:
        assert tgt_mask.shape[1] == T
    out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
    mask = torch.empty(B, T, device=dev, dtype=torch.uint8)
    lens = torch.empty(B, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pb_length_regulate(_lib.ptr(seq), _lib.ptr(dur), None, B, S, D, T, _lib.ptr(out), _lib.ptr(mask),
                                                  _lib.ptr(lens), _lib.stream_ptr(dev)))
    return out, (mask.bool() if tgt_mask is None else tgt_mask)


class Conv(nn.Module):
    """Parameter container for a channel-last Conv1d (reference duration.py:51-79)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)

    def forward(self, *a, **k):
        raise RuntimeError("Conv runs fused inside DurationPredictor.forward on the sm_100a kernels")


class DurationPredictor(nn.Module):
    def __init__(self, d_model, n_filter, kernel_size, dropout_p):
        super().__init__()
        self.d_model, self.n_filter, self.kernel_size = d_model, n_filter, kernel_size
        self.layers = nn.Sequential(
            Conv(d_model, n_filter, kernel_size, padding=(kernel_size - 1) // 2), nn.ReLU(), nn.LayerNorm(n_filter),
            nn.Dropout(dropout_p),
            Conv(n_filter, n_filter, kernel_size, padding=1), nn.ReLU(), nn.LayerNorm(n_filter), nn.Dropout(dropout_p))
        self.proj = nn.Linear(n_filter, 1)
        self._eng = TteEngine()

    def forward(self, x, mask=None):
        """log-durations f32 [B,S]; ``mask`` True = padded char -> 0.0 (reference duration.py:41-48)."""
        if self.training:
            raise _lib.ParrotB200Error("DurationPredictor is inference-only: call .eval() first")
        need_cuda(x, "DurationPredictor.forward")
        dev = x.device
        cfg = dict(src_vocab_size=1, src_pad_idx=0, n_speaker=0, d_model=self.d_model, n_head=2, n_enc_layer=0, n_dec_layer=0,
                   conv_n_filter=1024, conv_k1=9, conv_k2=1, dp_n_filter=self.n_filter, dp_kernel=self.kernel_size, n_codes=1,
                   max_len=2, fold_projections=1, gemm_mode=0)
        tensors = {"duration_predictor." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        hnd = self._eng.get(cfg, tensors, dev)
        B, S, _ = x.shape
        x = x.to(torch.float32).contiguous()
        src_mask = None if mask is None else (~mask).to(device=dev, dtype=torch.uint8).contiguous()
        out = torch.empty(B, S, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pb_tte_duration(hnd, _lib.ptr(x), _lib.ptr(src_mask), B, S, _lib.ptr(out), None,
                                                   _lib.stream_ptr(dev)))
        return out
