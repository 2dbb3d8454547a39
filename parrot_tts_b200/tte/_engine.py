"""Owns a pb_tte handle for a set of TTE weights on one device."""
import ctypes as C

import torch

from .. import _lib
from .._weights import WeightStamp


class TteEngine:
    def __init__(self):
        self.handle = None
        self.stamp = None

    def close(self):
        if self.handle is not None:
            _lib.load().pb_tte_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get(self, cfg_fields, tensors, device):
        """cfg_fields: dict for TteCfg; tensors: {checkpoint key: tensor} or a ``WeightStamp`` (cached tensor list: the
        per-call check then costs tens of microseconds instead of a state_dict walk).  Rebuilds when anything changed."""
        if isinstance(tensors, WeightStamp):
            wstamp = tensors.stamp()
        else:
            wstamp = tuple((k, v.data_ptr(), v._version, tuple(v.shape)) for k, v in tensors.items())
        stamp = (str(device), tuple(sorted(cfg_fields.items())), wstamp)
        if self.handle is not None and self.stamp == stamp:
            return self.handle
        if isinstance(tensors, WeightStamp):
            tensors = tensors.tensors()
        lib = _lib.load()
        self.close()
        cfg = _lib.TteCfg()
        for k, v in cfg_fields.items():
            setattr(cfg, k, int(v))
        with torch.cuda.device(device):
            hnd = C.c_void_p()
            _lib.check(lib.pb_tte_create(C.byref(cfg), C.byref(hnd)))
            try:
                for key, t in tensors.items():
                    src = t
                    t = t.detach().to(device=device, dtype=torch.float32).contiguous()
                    if t.data_ptr() != src.data_ptr():
                        # a conversion / copy ran on torch's current stream; set_weight copies on the legacy default stream, which
                        # does not order against non-blocking streams: wait for the temporary before handing its pointer over
                        torch.cuda.current_stream(device).synchronize()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.pb_tte_set_weight(hnd, key.encode(), _lib.ptr(t), shape, t.dim(), 0))
                _lib.check(lib.pb_tte_finalize(hnd, _lib.stream_ptr(device)))
            except Exception:
                lib.pb_tte_destroy(hnd)
                raise
        self.handle, self.stamp = hnd, stamp
        return hnd


    def _stamp_of(self, cfg_fields, wstamp, device):
        return (str(device), tuple(sorted(cfg_fields.items())), wstamp.stamp())

    def create_empty(self, cfg_fields, device):
        """Handle with the arena allocated and nothing uploaded (it is about to be received from another rank)."""
        lib = _lib.load()
        self.close()
        cfg = _lib.TteCfg()
        for k, v in cfg_fields.items():
            setattr(cfg, k, int(v))
        with torch.cuda.device(device):
            hnd = C.c_void_p()
            _lib.check(lib.pb_tte_create(C.byref(cfg), C.byref(hnd)))
        self.handle, self.stamp = hnd, None
        return hnd

    def adopt(self, cfg_fields, wstamp, device):
        _lib.check(_lib.load().pb_tte_mark_finalized(self.handle))
        self.stamp = self._stamp_of(cfg_fields, wstamp, device)


def need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.ParrotB200Error(f"{what} needs CUDA tensors: the TTE has no CPU path")
