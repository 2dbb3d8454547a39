"""Single-op entry points of libparrot_b200.so on torch CUDA tensors (used by the unit tests and by the
FFTBlock / DurationPredictor / length_regulator mirrors).  Every function runs the same sm_100a kernel
the fused model path uses; there is no fallback."""
import ctypes as C

import torch

from . import _lib


def _f32(t):
    return t.to(torch.float32).contiguous()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.ParrotB200Error("parrot_tts_b200.ops need CUDA tensors (no CPU path)")


def conv1d(x, w, bias=None, dilation=1, res=None, pre_slope=1.0, post_slope=1.0):
    """post(conv1d(pre(x), w, bias, 'same', dilation) [+ res]); x [B,Cin,L], w [Cout,Cin,K]."""
    _need_cuda(x, w, bias, res)
    x, w = _f32(x), _f32(w)
    B, Cin, L = x.shape
    Cout, _, K = w.shape
    bias = _f32(bias) if bias is not None else None
    res = _f32(res) if res is not None else None
    y = torch.empty(B, Cout, L, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(y),
                                         B, Cin, Cout, L, K, dilation, pre_slope, post_slope, _lib.stream_ptr(x.device)))
    return y


def conv_transpose1d(x, w, bias=None, stride=1, pre_slope=1.0):
    """conv_transpose1d(pre(x), w [Cin,Cout,K], bias, stride, padding=(K-stride)/2) -> [B,Cout,stride*L]."""
    _need_cuda(x, w, bias)
    x, w = _f32(x), _f32(w)
    B, Cin, L = x.shape
    _, Cout, K = w.shape
    bias = _f32(bias) if bias is not None else None
    y = torch.empty(B, Cout, L * stride, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_conv_transpose1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y),
                                                   B, Cin, Cout, L, K, stride, pre_slope, _lib.stream_ptr(x.device)))
    return y


def conv1d_tc(x, w, bias=None, dilation=1, res=None, pre_slope=1.0, post_slope=1.0, single_pass=False):
    """conv1d on the tensor-core path (split-bf16 tcgen05 implicit GEMM); same contract as conv1d."""
    _need_cuda(x, w, bias, res)
    x, w = _f32(x), _f32(w)
    B, Cin, L = x.shape
    Cout, _, K = w.shape
    bias = _f32(bias) if bias is not None else None
    res = _f32(res) if res is not None else None
    y = torch.empty(B, Cout, L, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_conv1d_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(y),
                                            B, Cin, Cout, L, K, dilation, pre_slope, post_slope, int(bool(single_pass)),
                                            _lib.stream_ptr(x.device)))
    return y


def conv_transpose1d_tc(x, w, bias=None, stride=1, pre_slope=1.0, single_pass=False):
    """conv_transpose1d on the tensor-core path (3-tap polyphase implicit GEMM); same contract as conv_transpose1d."""
    _need_cuda(x, w, bias)
    x, w = _f32(x), _f32(w)
    B, Cin, L = x.shape
    _, Cout, K = w.shape
    bias = _f32(bias) if bias is not None else None
    y = torch.empty(B, Cout, L * stride, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_conv_transpose1d_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y),
                                                      B, Cin, Cout, L, K, stride, pre_slope, int(bool(single_pass)),
                                                      _lib.stream_ptr(x.device)))
    return y


def wav_to_pcm16(wav):
    """(wav * 32768) -> int16 with the reference's cast semantics (utils/vocoder/inference.py:71-73)."""
    _need_cuda(wav)
    wav = _f32(wav)
    pcm = torch.empty(wav.shape, device=wav.device, dtype=torch.int16)
    if wav.numel() == 0:
        return pcm
    with torch.cuda.device(wav.device):
        _lib.check(_lib.load().pb_wav_to_pcm16(_lib.ptr(wav), wav.numel(), _lib.ptr(pcm), _lib.stream_ptr(wav.device)))
    return pcm


def to_host(t):
    """Device tensor -> numpy through a pinned staging tensor from torch's caching host allocator: a D2H copy into pageable
    memory (``.cpu()``) runs at ~10 GB/s, into pinned memory at PCIe speed - 0.3 ms of the 5 MB of PCM an 8 x 1024-unit batch
    produces.  The array keeps the pinned block alive; it returns to the cache when the array is dropped."""
    _need_cuda(t)
    t = t.contiguous()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def peak_normalize(pcm):
    """`librosa.util.normalize(audio.astype(np.float32))` on the device (utils/vocoder/inference.py:169, demo cell 13):
    int16 PCM [n] or [B, n] -> float32 of the same shape, every utterance divided by its own max |sample|
    (an all-zero utterance is returned unchanged)."""
    _need_cuda(pcm)
    if pcm.dtype != torch.int16:
        raise TypeError("peak_normalize expects the int16 PCM that generate() returns")
    pcm = pcm.contiguous()
    if pcm.dim() not in (1, 2):
        raise ValueError("peak_normalize: [n] or [B, n] expected")
    B, n = (1, pcm.numel()) if pcm.dim() == 1 else pcm.shape
    out = torch.empty(pcm.shape, device=pcm.device, dtype=torch.float32)
    peak = torch.empty(max(B, 1), device=pcm.device, dtype=torch.int32)
    if pcm.numel() == 0:
        return out
    with torch.cuda.device(pcm.device):
        _lib.check(_lib.load().pb_pcm16_peak_normalize(_lib.ptr(pcm), B, n, _lib.ptr(peak), _lib.ptr(out),
                                                        _lib.stream_ptr(pcm.device)))
    return out


def linear(x, w, bias=None, mode=0):
    """nn.Linear: x [M,K] . w[N,K]^T + bias (fp32)."""
    _need_cuda(x, w, bias)
    x, w = _f32(x), _f32(w)
    M, K = x.shape
    N = w.shape[0]
    bias = _f32(bias) if bias is not None else None
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_linear(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), M, N, K, mode, _lib.stream_ptr(x.device)))
    return y


def layernorm(x, gamma, beta):
    """nn.LayerNorm(256), eps 1e-5, over the last dim."""
    _need_cuda(x, gamma, beta)
    x = _f32(x)
    D = x.shape[-1]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pb_layernorm(_lib.ptr(x), _lib.ptr(_f32(gamma)), _lib.ptr(_f32(beta)), _lib.ptr(y),
                                            x.numel() // D, D, _lib.stream_ptr(x.device)))
    return y
