"""HiFi-GAN code generator with the reference's class names, constructor/forward signatures and
state_dict layout (reference utils/vocoder/models.py:13-169), executing on libparrot_b200.so.

The torch.nn conv modules below are *parameter containers only* - they give ``state_dict()`` /
``load_state_dict()`` exactly the reference's key names and shapes (old-style weight norm:
``weight_g`` / ``weight_v``; plain ``weight`` after ``remove_weight_norm()``).  ``forward`` never
calls them: it hands the tensors to the sm_100a kernels through the C ABI.  There is no PyTorch
or CPU fallback - calling ``forward`` on CPU tensors raises.
"""
import ctypes as C
import warnings

import torch
import torch.nn as nn

from .. import _lib
from .._weights import WeightStamp
from .utils import get_padding, init_weights

LRELU_SLOPE = 0.1


def _wn(module):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(module)


def _strip_wn(module):
    torch.nn.utils.remove_weight_norm(module)


class _ResBlockParams(nn.Module):
    """Parameter container shared by ResBlock1 / ResBlock2."""

    def _make(self, channels, kernel_size, dilations):
        convs = nn.ModuleList(
            _wn(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilations)
        convs.apply(init_weights)
        return convs

    def forward(self, x):
        raise RuntimeError("resblocks execute inside Generator.forward (fused sm_100a kernels); "
                           "use parrot_tts_b200.ops.conv1d for a single convolution")


class ResBlock1(_ResBlockParams):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = self._make(channels, kernel_size, dilation)
        self.convs2 = self._make(channels, kernel_size, [1] * len(dilation))

    def remove_weight_norm(self):
        for conv in list(self.convs1) + list(self.convs2):
            _strip_wn(conv)


class ResBlock2(_ResBlockParams):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = self._make(channels, kernel_size, dilation)

    def remove_weight_norm(self):
        for conv in self.convs:
            _strip_wn(conv)


class _Engine:
    """Owns one pb_voc handle for (module, device); rebuilt when the parameters change."""

    def __init__(self):
        self.handle = None
        self.stamp = None
        self.device = None

    def close(self):
        if self.handle is not None:
            _lib.load().pb_voc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Generator(nn.Module):
    #: pb_voc_cfg.conv_mode: 0 = fp32 SIMT kernels, 1 = tcgen05 split-bf16 (fp32-accurate, default), 2 = tcgen05 bf16
    conv_mode = 1

    def __init__(self, h):
        super().__init__()
        self.h = h
        if str(h.resblock) != "1":
            raise NotImplementedError("only resblock type '1' (the reference config) has sm_100a kernels")
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        c0 = h.upsample_initial_channel
        self.conv_pre = _wn(nn.Conv1d(getattr(h, "model_in_dim", 128), c0, 7, 1, padding=3))
        self.ups = nn.ModuleList(
            _wn(nn.ConvTranspose1d(c0 // 2 ** i, c0 // 2 ** (i + 1), k, u, padding=(k - u) // 2))
            for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList(
            ResBlock1(h, c0 // 2 ** (i + 1), k, d)
            for i in range(self.num_upsamples)
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes))
        self.conv_post = _wn(nn.Conv1d(c0 // 2 ** self.num_upsamples, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)
        self._eng = _Engine()
        self._wstamp = None
        self.hop = 1
        for u in h.upsample_rates:
            self.hop *= int(u)

    # ------------------------------------------------------------------ engine management
    def _has_code_path(self):
        return False

    def _cfg(self):
        h = self.h
        cfg = _lib.VocCfg()
        cfg.model_in_dim = int(getattr(h, "model_in_dim", 128))
        cfg.upsample_initial_channel = int(h.upsample_initial_channel)
        cfg.num_upsamples = self.num_upsamples
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cfg.upsample_rates[i] = int(u)
            cfg.upsample_kernel_sizes[i] = int(k)
        cfg.num_kernels = self.num_kernels
        dil = h.resblock_dilation_sizes
        cfg.num_dilations = len(dil[0])
        for j, k in enumerate(h.resblock_kernel_sizes):
            cfg.resblock_kernel_sizes[j] = int(k)
            if len(dil[j]) != cfg.num_dilations:
                raise ValueError("all resblocks must have the same number of dilations")
            for m, d in enumerate(dil[j]):
                cfg.resblock_dilation_sizes[j][m] = int(d)
        cfg.num_embeddings = 0
        cfg.embedding_dim = 0
        cfg.num_speakers = 0
        cfg.conv_mode = int(self.conv_mode)
        return cfg

    def _stamp(self, device):
        if self._wstamp is None:
            self._wstamp = WeightStamp(self)
        return (str(device), int(self.conv_mode)) + self._wstamp.stamp()

    def invalidate_weights(self):
        """Force the next forward to re-upload, re-fold and re-pack the weights.  Needed only after edits torch cannot see
        (``p.data.copy_()`` and the like bump no version counter); ``load_state_dict``, in-place ops, ``remove_weight_norm`` and
        ``.to()`` are detected."""
        if self._wstamp is None:
            self._wstamp = WeightStamp(self)
        self._wstamp.invalidate()

    def _engine(self, device):
        """pb_voc handle holding the folded + packed copy of the current parameters."""
        stamp = self._stamp(device)
        eng = self._eng
        if eng.handle is not None and eng.stamp == stamp:
            return eng.handle
        lib = _lib.load()
        eng.close()
        with torch.cuda.device(device):
            hnd = C.c_void_p()
            cfg = self._cfg()
            _lib.check(lib.pb_voc_create(C.byref(cfg), C.byref(hnd)))
            try:
                for key, t in self.state_dict().items():
                    src = t
                    t = t.detach().to(device=device, dtype=torch.float32).contiguous()
                    if t.data_ptr() != src.data_ptr():
                        # a conversion / copy ran on torch's current stream; set_weight copies on the legacy default stream, which
                        # does not order against non-blocking streams: wait for the temporary before handing its pointer over
                        torch.cuda.current_stream(device).synchronize()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.pb_voc_set_weight(hnd, key.encode(), _lib.ptr(t), shape, t.dim(), 0))
                _lib.check(lib.pb_voc_finalize(hnd, _lib.stream_ptr(device)))
            except Exception:
                lib.pb_voc_destroy(hnd)
                raise
        eng.handle, eng.stamp, eng.device = hnd, stamp, device
        return hnd

    # ---- packed-arena hand-over (sharding.broadcast_engines): the folded + packed weights live in ONE device arena whose layout is
    # a pure function of the config, so a rank that did not load the checkpoint receives the arena instead of the state dict
    def engine_arena(self, device, receive=False):
        """(device pointer, bytes) of the engine's weight arena.  ``receive=True`` creates an EMPTY engine (no upload, no fold) whose
        arena is about to be filled from another rank; call ``engine_adopt()`` once it has been."""
        lib = _lib.load()
        device = torch.device(device)
        if receive:
            self._eng.close()
            with torch.cuda.device(device):
                hnd = C.c_void_p()
                cfg = self._cfg()
                _lib.check(lib.pb_voc_create(C.byref(cfg), C.byref(hnd)))
            self._eng.handle, self._eng.stamp, self._eng.device = hnd, None, device
        else:
            hnd = self._engine(device)
        n = C.c_int64(0)
        _lib.check(lib.pb_voc_arena_bytes(hnd, C.byref(n)))
        return int(lib.pb_voc_arena_ptr(hnd)), int(n.value)

    def engine_adopt(self):
        """Declare the received arena complete: the engine now stands for this module's weights (its parameters are placeholders
        until something is loaded into them, which rebuilds the engine like any other weight change)."""
        _lib.check(_lib.load().pb_voc_mark_finalized(self._eng.handle))
        self._eng.stamp = self._stamp(self._eng.device)

    def _run(self, x=None, code=None, spkr=None, lens=None, pcm16=False):
        ref = x if x is not None else code
        if not ref.is_cuda:
            raise _lib.ParrotB200Error("Generator.forward needs CUDA tensors: the vocoder has no CPU path")
        device = ref.device
        lib = _lib.load()
        hnd = self._engine(device)
        B, T = (x.shape[0], x.shape[2]) if x is not None else code.shape
        # pcm16: conv_post's epilogue writes the int16 PCM itself and the f32 waveform never reaches HBM (the fp32 SIMT mode casts
        # from the waveform in a separate launch and needs it)
        want_wav = not pcm16 or int(self.conv_mode) == 0
        wav = torch.empty(B, 1, self.hop * T, device=device, dtype=torch.float32) if want_wav else None
        pcm = torch.empty(B, 1, self.hop * T, device=device, dtype=torch.int16) if pcm16 else None
        if x is not None:
            x = x.to(torch.float32).contiguous()
        else:
            code = code.to(torch.int64).contiguous()
            if spkr is not None:
                spkr = spkr.to(device=device, dtype=torch.int64).reshape(-1).contiguous()
                if spkr.numel() != B:
                    raise ValueError(f"spkr must hold one id per item, got {spkr.numel()} for batch {B}")
        if lens is not None:
            lens = torch.as_tensor(lens).to(device=device, dtype=torch.int32).contiguous()
        with torch.cuda.device(device):
            if pcm16:
                _lib.check(lib.pb_voc_forward_pcm16(hnd, _lib.ptr(x), _lib.ptr(code), _lib.ptr(spkr), _lib.ptr(lens),
                                                    B, T, _lib.ptr(wav), _lib.ptr(pcm), _lib.stream_ptr(device)))
            else:
                _lib.check(lib.pb_voc_forward(hnd, _lib.ptr(x), _lib.ptr(code), _lib.ptr(spkr), _lib.ptr(lens),
                                              B, T, _lib.ptr(wav), _lib.stream_ptr(device)))
        self.last_launches = int(lib.pb_voc_last_launches(hnd))
        return pcm if pcm16 else wav

    def check_ids(self):
        """Raise ``IndexError`` if a forward since the last check was given a unit / speaker id outside its embedding table at a
        live position, as the reference's ``nn.Embedding`` does (utils/vocoder/models.py:155-156).  The forward itself is
        asynchronous and clamps; this waits for the device."""
        if self._eng.handle is not None:
            with torch.cuda.device(self._eng.device):
                _lib.check(_lib.load().pb_voc_check_ids(self._eng.handle, _lib.stream_ptr(self._eng.device)))

    # ------------------------------------------------------------------ measurement hooks
    def set_timing(self, enable, device="cuda:0"):
        """Record CUDA events at the section boundaries of every forward (see pb_voc_set_timing)."""
        _lib.check(_lib.load().pb_voc_set_timing(self._engine(torch.device(device)), int(bool(enable))))

    def section_ms(self):
        """{section name: milliseconds} of the last timed forward (waits for it to finish)."""
        lib = _lib.load()
        n = lib.pb_voc_section_count(self._eng.handle)
        buf = (C.c_float * n)()
        _lib.check(lib.pb_voc_section_ms(self._eng.handle, buf, n))
        names = ["embed+conv_pre"]
        for i in range(self.num_upsamples):
            names += [f"ups{i}", f"resblocks{i}"]
        names.append("conv_post")
        return dict(zip(names, [float(v) for v in buf]))

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        """f32 [B, model_in_dim, T] -> f32 [B, 1, hop*T] (reference utils/vocoder/models.py:95-111)."""
        return self._run(x=x)

    def forward_pcm16(self, x):
        """forward + the generate() tail (x 32768, int16 cast; reference utils/vocoder/inference.py:71-73) as conv_post's
        epilogue -> i16 [B, 1, hop*T]: bitwise ``ops.wav_to_pcm16(self.forward(x))`` without the f32 waveform round trip."""
        return self._run(x=x, pcm16=True)

    def remove_weight_norm(self):
        for up in self.ups:
            _strip_wn(up)
        for rb in self.resblocks:
            rb.remove_weight_norm()
        _strip_wn(self.conv_pre)
        _strip_wn(self.conv_post)


class CodeGenerator(Generator):
    def __init__(self, h):
        super().__init__(h)
        self.dict = nn.Embedding(h.num_embeddings, h.embedding_dim)
        self.f0 = h.get("f0", None)
        self.multispkr = h.get("multispkr", None)
        if self.multispkr:
            self.spkr = nn.Embedding(10, h.embedding_dim)

    def _cfg(self):
        cfg = super()._cfg()
        cfg.num_embeddings = int(self.h.num_embeddings)
        cfg.embedding_dim = int(self.h.embedding_dim)
        cfg.num_speakers = int(self.spkr.num_embeddings) if self.multispkr else 0
        return cfg

    def forward(self, **kwargs):
        """code i64 [B,T], spkr i64 [B,1] -> f32 [B,1,hop*T] (reference utils/vocoder/models.py:153-169).

        ``lens`` (i32/i64 [B], optional, not in the reference) gives each item's true unit count: every
        item then equals its own batch-of-1 result instead of the padded-rectangle result.
        """
        extra = [k for k in kwargs if k not in ("code", "spkr", "f0", "lens", "pcm16")]
        if extra:
            raise NotImplementedError(f"extra conditioning features {extra} are not supported: model_in_dim is "
                                      "fully used by the unit and speaker embeddings")
        spkr = kwargs.get("spkr") if self.multispkr else None
        if self.multispkr and spkr is None:
            raise KeyError("spkr")
        return self._run(code=kwargs["code"], spkr=spkr, lens=kwargs.get("lens"), pcm16=bool(kwargs.get("pcm16", False)))

    def forward_pcm16(self, **kwargs):
        """``forward(**kwargs)`` with the int16 cast of generate() fused into conv_post -> i16 [B, 1, hop*T]."""
        return self.forward(**dict(kwargs, pcm16=True))
