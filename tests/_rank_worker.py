"""Worker of tests/test_multirank_gpu.py: one rank of an N-rank synthesis job (launched by torch.distributed.run).

Rank 0 alone owns the weights ("loads the checkpoint"); the packed weight arenas are broadcast once; every rank runs its
longest-processing-time-first share of the pool through serving.SpeechPipeline and writes its utterances' int16 PCM to
<outdir>/rank<r>.npz.  No collective on the data path."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def pool():
    g = torch.Generator().manual_seed(77)
    lens = [int(v) for v in torch.randint(3, 19, (14,), generator=g)]
    return [torch.randint(2, 150, (n,), generator=g) for n in lens], [i % 10 for i in range(len(lens))]


def models(load, dev):
    from parrot_tts_b200 import configs, synth
    from parrot_tts_b200.tte import Parrot
    from parrot_tts_b200.vocoder import CodeGenerator
    d = tempfile.mkdtemp(prefix="parrot_rank_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    cfg = configs.default_tte_config(root_path=d)
    h = configs.default_vocoder_config()
    voc, tte = CodeGenerator(h), Parrot(cfg, 150, 0)
    if load:
        voc.load_state_dict(synth.synth_vocoder_state_dict(h, seed=5))
        tte.load_state_dict(synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=5))
    return voc.to(dev), tte.eval().to(dev)


def main():
    outdir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n_dev = torch.cuda.device_count()
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]) % n_dev)
    torch.cuda.set_device(dev)
    # one rank per GPU over NCCL when the box has the GPUs; on a single-GPU box both ranks share it and the one broadcast goes over gloo
    if n_dev >= world:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    from parrot_tts_b200 import serving, sharding
    voc, tte = models(rank == 0, dev)
    nbytes = sharding.broadcast_engines([voc, tte], src=0, device=dev)
    phones, spk = pool()
    unit_lens = [8 * int(p.numel()) for p in phones]                 # cost estimate only (the share must not depend on results)
    buckets = serving.plan_shares(unit_lens, world, max_batch=3, max_pad_ratio=1.25)[rank]
    mine = sorted(i for b in buckets for i in b)
    pipe = serving.SpeechPipeline(tte, voc, max_batch=3)
    audio = pipe(phones, spk, batches=buckets)
    assert [a is not None for a in audio] == [i in mine for i in range(len(phones))]
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **{f"u{i}": audio[i] for i in mine}, broadcast_bytes=np.int64(nbytes))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
