"""Data-parallel utterance sharding over the GPUs of one box.

Replaces the reference's ``multiprocessing.Pool(8)`` + queue of GPU ids in which every worker loads the
checkpoint from disk (reference utils/vocoder/inference.py:201-205,93-109,253-256): here one process per GPU
(torchrun), rank 0 owns the checkpoint, ONE broadcast of the packed weights at init (NCCL on GPUs, gloo in the
CPU tests), utterances assigned by longest-processing-time-first on their cost, and no collective on the
per-utterance path.
"""
import torch
import torch.distributed as dist


def utterance_cost(n_units, with_tte=True):
    """FLOP estimate per utterance: vocoder 321.66 MFLOP/unit; TTE decoder
    4*(6291456 + 1024*T) + 512000 FLOP per unit at T keys (attention makes it quadratic)."""
    n = float(n_units)
    c = 321656832.0 * n
    if with_tte:
        c += n * (4.0 * (6291456.0 + 1024.0 * n) + 512000.0)
    return c


def lpt_assign(costs, world_size):
    """Greedy longest-processing-time-first: -> list (per rank) of utterance indices.  Deterministic: ties are
    broken by index, so every rank computes the same assignment without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    bins = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += costs[i]
    return bins


def length_buckets(indices, lengths, max_batch, max_pad_ratio=1.25):
    """Group one rank's utterances (longest first) into batches whose shortest item is within ``max_pad_ratio`` of
    the longest, at most ``max_batch`` items each: bounds the padding the rectangular kernels compute."""
    idx = sorted(indices, key=lambda i: (-lengths[i], i))
    out, cur = [], []
    for i in idx:
        if cur and (len(cur) >= max_batch or lengths[cur[0]] > max_pad_ratio * lengths[i]):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def pack_state_dict(sd):
    """{key: tensor} -> (flat f32 tensor, manifest [(key, shape)]) in sorted-key order."""
    keys = sorted(sd)
    manifest = [(k, tuple(sd[k].shape)) for k in keys]
    flat = torch.cat([sd[k].detach().to(torch.float32).reshape(-1).cpu() for k in keys]) if keys else torch.zeros(0)
    return flat, manifest


def unpack_state_dict(flat, manifest):
    sd, off = {}, 0
    for k, shape in manifest:
        n = 1
        for s in shape:
            n *= s
        sd[k] = flat[off:off + n].reshape(shape).clone()
        off += n
    return sd


def broadcast_state_dict(sd, src=0, device=None, group=None):
    """One broadcast of all weights from ``src``: returns the state dict on every rank (``sd`` may be None off-src).
    The manifest (names + shapes, a few KB) travels as an object; the payload is a single flat tensor."""
    rank = dist.get_rank(group)
    meta = [None]
    flat = None
    if rank == src:
        flat, manifest = pack_state_dict(sd)
        meta = [manifest]
    dist.broadcast_object_list(meta, src=src, group=group)
    manifest = meta[0]
    total = 0
    for _, shape in manifest:
        n = 1
        for s in shape:
            n *= s
        total += n
    if rank != src:
        flat = torch.empty(total, dtype=torch.float32)
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src, group=group)
    return unpack_state_dict(flat, manifest)


class _DeviceBytes:
    """Raw device memory as a ``__cuda_array_interface__`` object (torch wraps it without copying)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def broadcast_engines(modules, src=0, device=None, group=None):
    """ONE broadcast per model of the PACKED weight arena: ``src`` has loaded the checkpoint into its modules;
    its engines upload, fold (weight norm, projection folding) and pack once, and every other rank receives the finished arena
    straight into its own engine (``pb_*_arena_ptr``) - no state dict travels, nobody else folds or packs.  ``modules``: the
    ``Parrot`` / ``CodeGenerator`` instances of this rank (constructed from the same configs on every rank; on the other ranks their
    parameters stay placeholders).  Replaces every worker's own ``torch.load`` of the checkpoint (reference
    utils/vocoder/inference.py:104-109)."""
    rank = dist.get_rank(group)
    device = torch.device(device if device is not None else "cuda")
    if device.index is None:      # the engines key their device copy by the full device string
        device = torch.device("cuda", torch.cuda.current_device())
    total = 0
    for m in modules:
        ptr, nbytes = m.engine_arena(device, receive=(rank != src))
        dist.broadcast(torch.as_tensor(_DeviceBytes(ptr, nbytes), device=device), src=src, group=group)
        if rank != src:
            m.engine_adopt()
        total += nbytes
    torch.cuda.synchronize(device)
    return total
