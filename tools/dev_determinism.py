"""Dev: run-to-run bitwise determinism of the vocoder forward and the TTE logits (a data race in the pipelines would show up here)."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator
from parrot_tts_b200.tte import Parrot
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
h = configs.default_vocoder_config()
gen = CodeGenerator(h); gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); gen = gen.cuda()
code = torch.randint(0, 1000, (8, 1024)).cuda(); spkr = torch.randint(0, 10, (8, 1)).cuda()
ref = gen(code=code, spkr=spkr).clone()
bad = 0
for i in range(N):
    w = gen(code=code, spkr=spkr)
    bad += int(not torch.equal(w, ref))
print("vocoder 8x1024:", N, "repeats,", bad, "differ; max|wav| =", float(ref.abs().max()))
d = tempfile.mkdtemp(); json.dump({f"s{i}": i for i in range(10)}, open(os.path.join(d, "speakers.json"), "w"))
cfg = configs.default_tte_config(root_path=d)
m = Parrot(cfg, 150, 0); m.load_state_dict(synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8")); m = m.eval().cuda()
ph = torch.randint(2, 150, (32, 256)).cuda()
batch = {"phones": ph, "src_mask": ph != 0, "speaker": (torch.arange(32) % 10).cuda()}
ref = m(batch, inference=True)[0].clone()
bad = 0
for i in range(N):
    lg = m(batch, inference=True)[0]
    bad += int(not torch.equal(lg, ref))
print("TTE 32x256 -> 2048 logits:", N, "repeats,", bad, "differ")
