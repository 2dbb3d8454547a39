"""FFT block pieces with the reference's class names and parameter layout (reference modules/fft.py),
executing on libparrot_b200.so.  The torch.nn members are parameter containers; no forward below calls them."""
import ctypes as C
import math

import torch
from torch import nn

from .. import _lib
from ._engine import TteEngine, need_cuda


class SinusoidalPosEmb(nn.Module):
    def __init__(self, max_len, d_model):
        super().__init__()
        self.max_len = max_len
        self.d_model = d_model
        self.register_buffer("pe", self.positionalencoding1d(d_model, max_len))

    @staticmethod
    def positionalencoding1d(d_model, length):
        """sin on even / cos on odd columns, angular step exp(-ln(1e4) * 2i / d_model) (reference fft.py:22-38)."""
        if d_model % 2:
            raise ValueError(f"Cannot use sin/cos positional encoding with odd dim (got dim={d_model})")
        pos = torch.arange(0, length).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
        pe = torch.zeros(length, d_model)
        pe[:, 0::2] = torch.sin(pos * freq)
        pe[:, 1::2] = torch.cos(pos * freq)
        return pe

    def forward(self, x):
        """x + pe[x.size(1)] - the reference adds the single row indexed by the sequence LENGTH to every
        position (fft.py:17-19) and raises IndexError when the length reaches max_len."""
        need_cuda(x, "SinusoidalPosEmb.forward")
        B, L, D = x.shape
        if L >= self.max_len:
            raise IndexError(f"index {L} is out of bounds for dimension 0 with size {self.max_len}")
        # one-to-one "regulation" (every duration = 1) with the row add fused: same kernel the decoder input uses
        x = x.to(torch.float32).contiguous()
        dur = torch.ones(B, L, dtype=torch.int64, device=x.device)
        out = torch.empty_like(x)
        lens = torch.empty(B, dtype=torch.int64, device=x.device)
        row = self.pe[L].to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pb_length_regulate(_lib.ptr(x), _lib.ptr(dur), _lib.ptr(row), B, L, D, L, _lib.ptr(out), None,
                                                      _lib.ptr(lens), _lib.stream_ptr(x.device)))
        return out


class Attention(nn.Module):
    """Parameter container: qkv Linear + nn.MultiheadAttention + wo Linear, all bias-free (reference fft.py:41-51)."""

    def __init__(self, d_model, n_head, dropout_p=0.1):
        super().__init__()
        assert d_model % n_head == 0
        self.d_model, self.n_head, self.d_head = d_model, n_head, d_model // n_head
        self.qkv = nn.Linear(d_model, 3 * d_model, bias=False)
        self.mha = nn.MultiheadAttention(d_model, n_head, dropout=dropout_p, bias=False, batch_first=True)
        self.wo = nn.Linear(d_model, d_model, bias=False)
        self.dropout_p = dropout_p

    def forward(self, *a, **k):
        raise RuntimeError("Attention runs fused inside FFTBlock.forward on the sm_100a kernels")


class ConvLayer(nn.Module):
    """Parameter container: Conv1d(d, n_filter, k0) -> ReLU -> Conv1d(n_filter, d, k1) (reference fft.py:62-76)."""

    def __init__(self, d_model, n_filter, kernel_sizes):
        super().__init__()
        self.conv1 = nn.Conv1d(d_model, n_filter, kernel_sizes[0], padding=(kernel_sizes[0] - 1) // 2)
        self.conv2 = nn.Conv1d(n_filter, d_model, kernel_sizes[1], padding=(kernel_sizes[1] - 1) // 2)

    def forward(self, *a, **k):
        raise RuntimeError("ConvLayer runs fused inside FFTBlock.forward on the sm_100a kernels")


class FFTBlock(nn.Module):
    def __init__(self, d_model, n_head, n_filter, kernel_sizes, dropout_p):
        super().__init__()
        self.d_model, self.n_head, self.n_filter, self.kernel_sizes = d_model, n_head, n_filter, list(kernel_sizes)
        self.attention = Attention(d_model, n_head, dropout_p)
        self.convlayer = ConvLayer(d_model, n_filter, kernel_sizes)
        self.attn_norm = nn.LayerNorm(d_model)
        self.conv_norm = nn.LayerNorm(d_model)
        self._eng = TteEngine()
        self.fold_projections = True
        self.gemm_mode = 0          # pb_tte_cfg.gemm_mode: 0 = fp32 SIMT, 1 = tcgen05 split-fp16

    def forward(self, x, key_padding_mask=None, attn_mask=None):
        """h = x + Attn(LN(x)); out = h + Conv(LN(h))  (reference fft.py:94-100).
        key_padding_mask: bool [B, L], True = key ignored.  attn_mask is not supported (the reference never passes one)."""
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is not supported")
        if self.training and (self.attention.dropout_p > 0):
            raise _lib.ParrotB200Error("FFTBlock is inference-only: call .eval() first (dropout is not implemented)")
        need_cuda(x, "FFTBlock.forward")
        dev = x.device
        cfg = dict(src_vocab_size=1, src_pad_idx=0, n_speaker=0, d_model=self.d_model, n_head=self.n_head, n_enc_layer=1,
                   n_dec_layer=0, conv_n_filter=self.n_filter, conv_k1=self.kernel_sizes[0], conv_k2=self.kernel_sizes[1],
                   dp_n_filter=256, dp_kernel=3, n_codes=1, max_len=2, fold_projections=int(self.fold_projections), gemm_mode=int(self.gemm_mode))
        tensors = {"encoder_layers.0." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        hnd = self._eng.get(cfg, tensors, dev)
        B, L, _ = x.shape
        x = x.to(torch.float32).contiguous()
        valid = None
        if key_padding_mask is not None:
            valid = (~key_padding_mask).to(device=dev, dtype=torch.uint8).contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pb_tte_fft_block(hnd, 0, 0, _lib.ptr(x), _lib.ptr(valid), B, L, _lib.ptr(y), _lib.stream_ptr(dev)))
        return y
