"""Dev: per-section time of the tensor-core vocoder forward under PB_TC_ABLATE settings (set in the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator
h = configs.default_vocoder_config()
gen = CodeGenerator(h); gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); gen = gen.to("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8; T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
code = torch.randint(0, 1000, (B, T), device="cuda:0"); spkr = torch.randint(0, 10, (B, 1), device="cuda:0")
for _ in range(3): gen(code=code, spkr=spkr)
gen.set_timing(True)
acc = {}
for _ in range(5):
    gen(code=code, spkr=spkr)
    for k, v in gen.section_ms().items(): acc[k] = acc.get(k, 0) + v / 5
print(B, T, os.environ.get("PB_TC_ABLATE", "0"), "total %.2f" % sum(acc.values()), {k: round(v, 2) for k, v in acc.items() if k.startswith("res")})
