"""Dev: time Parrot.infer_device at cfg3 (B=32, S=256 -> T=2048) per gemm_mode, and compare ids between modes."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.tte import Parrot
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = tempfile.mkdtemp()
json.dump({f"s{i}": i for i in range(10)}, open(os.path.join(d, "speakers.json"), "w"))
cfg = configs.default_tte_config(root_path=d)
sd = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8")
g = torch.Generator().manual_seed(0)
ph = torch.randint(2, 150, (B, S), generator=g).cuda()
batch = {"phones": ph, "src_mask": ph != 0, "speaker": (torch.arange(B) % 10).cuda()}
out = {}
for mode in (0, 1):
    m = Parrot(cfg, 150, 0); m.load_state_dict(sd); m = m.eval().cuda(); m.gemm_mode = mode
    for _ in range(2): ids, n = m.infer_device(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ids, n = m.infer_device(batch)
    e1.record(); torch.cuda.synchronize()
    out[mode] = ids.cpu()
    print(f"mode {mode}: {e0.elapsed_time(e1)/3:.2f} ms per batch  T={ids.shape[1]} launches={m.last_launches}")
print("id mismatches mode1 vs mode0:", int((out[0] != out[1]).sum()), "of", out[0].numel())
