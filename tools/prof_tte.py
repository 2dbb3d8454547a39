"""Dev/profiling driver: one warm-up + ONE Parrot.infer_device at cfg3 (B=32, S=256 -> T=2048) for ncu -s/-c selection."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.tte import Parrot
B, S = 32, 256
d = tempfile.mkdtemp()
json.dump({f"s{i}": i for i in range(10)}, open(os.path.join(d, "speakers.json"), "w"))
cfg = configs.default_tte_config(root_path=d)
m = Parrot(cfg, 150, 0); m.load_state_dict(synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8")); m = m.eval().cuda()
ph = torch.randint(2, 150, (B, S), generator=torch.Generator().manual_seed(0)).cuda()
batch = {"phones": ph, "src_mask": ph != 0, "speaker": (torch.arange(B) % 10).cuda()}
for _ in range(2):
    ids, n = m.infer_device(batch)
torch.cuda.synchronize()
print("done", ids.shape, m.last_launches)
