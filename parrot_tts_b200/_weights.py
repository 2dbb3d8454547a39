"""Cheap change detection for a module's weights.

Every forward must know whether the device copy (folded + packed by ``*_finalize``) still matches the module's parameters.
Walking ``state_dict()`` per call costs ~0.5 ms for the vocoder's 295 tensors - more than the launches of a small forward -
so the tensor list is cached and only (data_ptr, version counter, shape) of each tensor is compared per call (~60 us).  The cached
LIST is rebuilt whenever ANY module registers a parameter or buffer (a process-wide counter fed by torch's registration hooks:
``remove_weight_norm``, ``load_state_dict(assign=True)``, manual re-registration) - the counter itself is not part of the
comparison key, so unrelated registrations do not make an engine look stale; in-place updates bump the tensors' version
counters and ``.to()`` / ``.cuda()`` change their data pointers, so both are seen by the per-call comparison."""
from torch.nn.modules import module as _tm

_structure = [0]


def _bump(*_args, **_kwargs):
    _structure[0] += 1
    return None


_tm.register_module_parameter_registration_hook(_bump)
_tm.register_module_buffer_registration_hook(_bump)


class WeightStamp:
    def __init__(self, module, prefix=""):
        self.module, self.prefix = module, prefix
        self._seen = -1
        self._keys, self._tensors = (), ()

    def _refresh(self):
        if self._seen != _structure[0]:
            sd = self.module.state_dict(keep_vars=True)
            self._keys = tuple(self.prefix + k for k in sd)
            self._tensors = tuple(sd.values())
            self._seen = _structure[0]

    def tensors(self):
        """{checkpoint key: tensor} of the module as it is now."""
        self._refresh()
        return dict(zip(self._keys, self._tensors))

    def stamp(self):
        """Comparison key of the module's weights AS THEY ARE NOW: names, storage, version counters and shapes.  The process-wide
        registration counter only decides when the cached tensor list is rebuilt - it is not part of the key, so registering
        parameters elsewhere in the process (a second model, an unrelated ``nn.Linear``) leaves this module's engine alone.
        Edits through ``.data`` do not bump version counters: call ``invalidate()`` after them."""
        self._refresh()
        return (self._epoch, self._keys) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in self._tensors)

    _epoch = 0

    def invalidate(self):
        """Force the next forward to re-upload (after in-place edits torch cannot see, e.g. ``p.data.copy_()``)."""
        self._epoch += 1
