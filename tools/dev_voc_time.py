"""Dev helper: time CodeGenerator.forward on the GPU (CUDA events, L2-exceeding inputs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
h = configs.default_vocoder_config()
gen = CodeGenerator(h)
gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0))
gen = gen.to("cuda:0")
code = torch.randint(0, 1000, (B, T), device="cuda:0")
spkr = torch.randint(0, 10, (B, 1), device="cuda:0")
for _ in range(3):
    wav = gen(code=code, spkr=spkr)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record()
    wav = gen(code=code, spkr=spkr)
ev[5].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
ms = min(ts)
units = B * T
flop = 321656832.0 * units
print(f"B={B} T={T} ms/forward={ms:.3f} (all {['%.2f' % t for t in ts]}) samples/s={units*320/ms*1e3:.4e} "
      f"RTF16k={units*320/ms*1e3/16000:.0f} TFLOP/s={flop/ms/1e9:.2f} launches={gen.last_launches} "
      f"HBM-alg GB/s={3616000.0*units/ms/1e6:.0f}")
