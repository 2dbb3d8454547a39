"""N-rank synthesis == 1-rank synthesis, bitwise, on the device: rank 0 owns the weights, ONE broadcast of the
packed arenas, LPT shares, no collective on the data path.  Replaces the reference's 8-process pool in which every worker loads
the checkpoint itself (utils/vocoder/inference.py:201-205,253-256)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_produce_the_single_rank_pcm_bitwise(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _rank_worker as w
    from parrot_tts_b200 import serving
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29671", os.path.join(ROOT, "tests", "_rank_worker.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    voc, tte = w.models(True, torch.device("cuda:0"))
    phones, spk = w.pool()
    want = serving.SpeechPipeline(tte, voc, max_batch=3)(phones, spk)              # the whole pool on one rank, in this process
    got, seen = {}, []
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert int(z["broadcast_bytes"]) > 100 << 20                                 # both packed arenas travelled (~0.2 GB)
        for k in z.files:
            if k.startswith("u"):
                got[int(k[1:])] = z[k]
                seen.append(int(k[1:]))
    assert sorted(seen) == list(range(len(phones)))                                   # every utterance exactly once
    for i, a in enumerate(want):
        assert a.dtype == np.int16 and np.array_equal(got[i], a), i
