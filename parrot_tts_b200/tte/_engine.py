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
                    t = t.detach().to(device=device, dtype=torch.float32).contiguous()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.pb_tte_set_weight(hnd, key.encode(), _lib.ptr(t), shape, t.dim(), 0))
                _lib.check(lib.pb_tte_finalize(hnd, _lib.stream_ptr(device)))
            except Exception:
                lib.pb_tte_destroy(hnd)
                raise
        self.handle, self.stamp = hnd, stamp
        return hnd


def need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.ParrotB200Error(f"{what} needs CUDA tensors: the TTE has no CPU path")
