"""Dev/profiling driver: warm-up forward(s) then ONE CodeGenerator forward at B x T (for ncu -s/-c selection).
usage: prof_voc.py [B] [T] [mode] [warmups]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = "cuda:0"
h = configs.default_vocoder_config()
gen = CodeGenerator(h); gen.conv_mode = mode
gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); gen = gen.to(dev)
code = torch.randint(0, 1000, (B, T), device=dev); spkr = torch.randint(0, 10, (B, 1), device=dev)
for _ in range(warm + 1):
    wav = gen(code=code, spkr=spkr)
torch.cuda.synchronize()
print("done", float(wav.abs().max()))
