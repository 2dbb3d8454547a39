"""Multi-GPU host logic on CPU: LPT assignment, length bucketing, and the one-off weight broadcast over a
world_size-2 gloo group (the NCCL path runs the same code on device tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parrot_tts_b200 import configs, sharding, synth


def test_lpt_covers_everything_and_balances():
    lens = np.random.default_rng(0).integers(256, 4097, 256).tolist()     # BASELINE config 4 lengths
    costs = [sharding.utterance_cost(n) for n in lens]
    bins = sharding.lpt_assign(costs, 8)
    flat = sorted(i for b in bins for i in b)
    assert flat == list(range(256))
    load = [sum(costs[i] for i in b) for b in bins]
    assert max(load) / (sum(load) / 8) < 1.02           # LPT: within 2 % of perfect balance here
    assert bins == sharding.lpt_assign(costs, 8)        # deterministic -> no communication needed


def test_lpt_edge_cases():
    assert sharding.lpt_assign([], 4) == [[], [], [], []]
    assert sharding.lpt_assign([5.0], 2) == [[0], []]
    assert sharding.lpt_assign([1.0, 1.0, 1.0], 1) == [[0, 1, 2]]


def test_length_buckets_bound_padding():
    lens = np.random.default_rng(1).integers(256, 4097, 64).tolist()
    buckets = sharding.length_buckets(list(range(64)), lens, max_batch=8, max_pad_ratio=1.25)
    assert sorted(i for b in buckets for i in b) == list(range(64))
    for b in buckets:
        assert len(b) <= 8
        assert max(lens[i] for i in b) <= 1.25 * min(lens[i] for i in b)


def test_pack_unpack_roundtrip():
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=3)
    flat, manifest = sharding.pack_state_dict(sd)
    back = sharding.unpack_state_dict(flat, manifest)
    assert set(back) == set(sd)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = configs.default_vocoder_config()
        sd = synth.synth_vocoder_state_dict(h, seed=5) if rank == 0 else None
        got = sharding.broadcast_state_dict(sd, src=0)
        ref = synth.synth_vocoder_state_dict(h, seed=5)
        ok = set(got) == set(ref) and all(torch.equal(got[k], ref[k]) for k in ref)
        lens = np.random.default_rng(0).integers(256, 4097, 32).tolist()
        mine = sharding.lpt_assign([sharding.utterance_cost(n) for n in lens], world)[rank]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        q.put((rank, ok, gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_broadcast_and_sharding_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, gathered in res:
        assert ok, f"rank {rank}: broadcast weights differ"
        assert sorted(i for b in gathered for i in b) == list(range(32))    # shards are disjoint and complete


def test_stream_chunks_cover_the_utterance():
    """Chunk plan of the long-form streaming path (BASELINE configs[4]): kept ranges tile [0, T), windows add the halo."""
    from parrot_tts_b200.vocoder.inference import stream_chunks
    for T, chunk, halo in [(16384, 1024, 32), (700, 256, 32), (5, 1024, 32), (1025, 1024, 0)]:
        plan = stream_chunks(T, chunk, halo)
        assert plan[0][2] == 0 and plan[-1][3] == T
        for (lo, hi, a, b), nxt in zip(plan, plan[1:] + [None]):
            assert 0 <= lo <= a < b <= hi <= T and b - a <= chunk
            assert lo == max(a - halo, 0) and hi == min(b + halo, T)
            if nxt is not None:
                assert nxt[2] == b


def test_serving_plan_is_a_partition_with_bounded_padding():
    """Batching front door (parrot_tts_b200.serving): every utterance lands in exactly one place, batches respect the
    size and padding bounds, long utterances are streamed, and the ranks' shares partition the request."""
    import numpy as np
    from parrot_tts_b200 import serving
    lengths = np.random.default_rng(0).integers(0, 12000, 300).tolist()
    lengths[5] = 0
    batches, long_form, empty = serving.plan_batches(lengths, max_batch=8, max_pad_ratio=1.25, long_form_units=8192)
    seen = sorted([i for b in batches for i in b] + long_form + empty)
    assert seen == list(range(len(lengths)))
    assert all(lengths[i] > 8192 for i in long_form) and all(lengths[i] == 0 for i in empty)
    for b in batches:
        assert 1 <= len(b) <= 8
        assert lengths[b[0]] == max(lengths[i] for i in b)
        assert lengths[b[0]] <= 1.25 * min(lengths[i] for i in b)
    shares = [serving.rank_share(lengths, 4, r) for r in range(4)]
    assert sorted(i for s in shares for i in s) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in s) for s in shares]
    assert max(loads) <= 1.02 * (sum(loads) / 4) + max(lengths)          # LPT: balanced to within one utterance
    sub_batches, sub_long, sub_empty = serving.plan_batches(lengths, indices=shares[1], long_form_units=8192)
    assert sorted([i for b in sub_batches for i in b] + sub_long + sub_empty) == shares[1]


def test_scheduling_invariants_property_based():
    """Property tests (hypothesis) of the host scheduling logic: LPT shares partition the request and are balanced to within
    the largest item; length buckets respect the size / padding bounds; stream chunks tile the utterance exactly once and
    every window carries the requested halo (clipped at the ends)."""
    from hypothesis import given, settings, strategies as st
    from parrot_tts_b200 import serving, sharding
    from parrot_tts_b200.vocoder import stream_chunks

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(0, 5000), min_size=0, max_size=80), st.integers(1, 8), st.integers(1, 9),
           st.floats(1.0, 2.0))
    def sched(lengths, world, max_batch, ratio):
        costs = [sharding.utterance_cost(n, with_tte=False) for n in lengths]
        bins = sharding.lpt_assign(costs, world)
        assert sorted(i for b in bins for i in b) == list(range(len(lengths)))
        if lengths:
            loads = [sum(costs[i] for i in b) for b in bins]
            assert max(loads) - min(loads) <= max(costs) + 1e-6
        for r in range(world):
            batches, long_form, empty = serving.plan_batches(lengths, max_batch, ratio, 4096, indices=bins[r])
            assert sorted([i for b in batches for i in b] + long_form + empty) == sorted(bins[r])
            for b in batches:
                assert 1 <= len(b) <= max_batch
                assert all(0 < lengths[i] <= 4096 for i in b)
                assert lengths[b[0]] <= ratio * min(lengths[i] for i in b) + 1e-9

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 20000), st.integers(1, 2048), st.integers(0, 64))
    def chunks(n, chunk, halo):
        cs = stream_chunks(n, chunk, halo)
        assert cs[0][2] == 0 and cs[-1][3] == n
        for (lo, hi, a, b), nxt in zip(cs, cs[1:] + [None]):
            assert 0 <= lo <= a < b <= hi <= n and b - a <= chunk
            assert lo == max(a - halo, 0) and hi == min(b + halo, n)
            if nxt is not None:
                assert nxt[2] == b

    sched()
    chunks()


def test_chunk_halo_is_sufficient_on_the_oracle():
    """The claim the streaming path rests on, checked on the CPU oracle (the reference graph): synthesising the windows of
    ``stream_chunks`` and keeping each chunk's own samples equals the unchunked waveform when halo >= the receptive-field radius
    (~23 units) - and does not with a short halo, so the test can see a broken plan."""
    import torch
    from oracle import voc_oracle
    from parrot_tts_b200 import configs, synth
    from parrot_tts_b200.vocoder.inference import stream_chunks
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=0)
    g = torch.Generator().manual_seed(3)
    T, hop = 150, 320
    code = torch.randint(0, 1000, (1, T), generator=g)
    spkr = torch.tensor([[4]])
    with torch.no_grad():
        full = voc_oracle.code_generator_forward(sd, h, code, spkr)[0, 0]

        def chunked(halo):
            parts = []
            for lo, hi, a, b in stream_chunks(T, 50, halo):
                w = voc_oracle.code_generator_forward(sd, h, code[:, lo:hi], spkr)[0, 0]
                parts.append(w[(a - lo) * hop:(b - lo) * hop])
            return torch.cat(parts)

        ok, short = chunked(24), chunked(6)
    assert ok.shape == full.shape == (T * hop,)
    assert float((ok - full).abs().max()) < 1e-5          # same arithmetic, another summation order inside the CPU conv (~1e-6)
    assert float((short - full).abs().max()) > 1e-4


def test_receptive_field_radius_on_the_oracle():
    """One changed unit moves the waveform only within +-24 units of it (what `halo_units` has to cover)."""
    import torch
    from oracle import voc_oracle
    from parrot_tts_b200 import configs, synth
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=0)
    g = torch.Generator().manual_seed(5)
    T, hop, mid = 96, 320, 48
    code = torch.randint(0, 1000, (1, T), generator=g)
    other = code.clone()
    other[0, mid] = (other[0, mid] + 500) % 1000
    spkr = torch.tensor([[0]])
    with torch.no_grad():
        d = (voc_oracle.code_generator_forward(sd, h, code, spkr) - voc_oracle.code_generator_forward(sd, h, other, spkr))[0, 0].abs()
    touched = torch.nonzero(d > 1e-7).flatten()
    lo_u, hi_u = int(touched.min()) // hop, int(touched.max()) // hop
    assert mid - 24 <= lo_u <= mid and mid <= hi_u <= mid + 24, (lo_u, hi_u)
    assert hi_u - lo_u >= 20                               # and it really is that wide: a halo of a few units would not do


def test_plan_shares_buckets_globally_then_balances():
    """serving.plan_shares: the pool is bucketed by length FIRST (full, tightly padded buckets whatever the number of ranks), then the
    buckets are spread longest-processing-time-first.  A partition of the pool, identical on every rank, padding and imbalance bounded
    on the BASELINE configs[3] pool - and better than splitting utterance by utterance and bucketing per rank."""
    from parrot_tts_b200 import serving
    lengths = [8 * ((int(v) + 7) // 8) for v in np.random.default_rng(0).integers(256, 4097, 256)]
    cost = lambda n: sharding.utterance_cost(n, with_tte=True)
    ideal = sum(cost(n) for n in lengths)
    for world in (1, 2, 4, 8):
        shares = serving.plan_shares(lengths, world, max_batch=8, max_pad_ratio=1.1)
        assert shares == serving.plan_shares(lengths, world, max_batch=8, max_pad_ratio=1.1)        # deterministic: no communication needed
        flat = sorted(i for rank in shares for b in rank for i in b)
        assert flat == list(range(256))
        loads = []
        for rank in shares:
            for b in rank:
                assert 1 <= len(b) <= 8 and lengths[b[0]] == max(lengths[i] for i in b) and lengths[b[0]] <= 1.1 * min(lengths[i] for i in b)
            loads.append(sum(len(b) * cost(lengths[b[0]]) for b in rank))
        assert max(loads) <= 1.05 * ideal / world, (world, max(loads) / (ideal / world))
        # the utterance-by-utterance split pays more padded work on its slowest rank
        bins = sharding.lpt_assign([cost(n) for n in lengths], world)
        per_rank = [sum(len(b) * cost(lengths[b[0]]) for b in sharding.length_buckets(bins[r], lengths, 8, 1.25)) for r in range(world)]
        assert max(loads) <= max(per_rank)
    assert serving.plan_shares([0, 5, 0], 2) == [[[1]], []]                      # empty utterances are nobody's work
