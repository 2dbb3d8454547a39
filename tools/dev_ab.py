"""Dev: one A/B data point in a few seconds - time CodeGenerator.forward (8 x 1024 units unless given) under whatever switches the
environment sets, and print a checksum of the waveform so that two runs can be compared bitwise.   python tools/dev_ab.py [B T]"""
import hashlib, os, sys
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
g = torch.Generator().manual_seed(0)
code = torch.randint(0, 1000, (B, T), generator=g).cuda()
spkr = torch.randint(0, 10, (B, 1), generator=g).cuda()
wav = gen(code=code, spkr=spkr)
torch.cuda.synchronize()
digest = hashlib.sha1(wav.cpu().numpy().tobytes()).hexdigest()[:16]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
for i in range(6):
    ev[i].record()
    wav = gen(code=code, spkr=spkr)
ev[6].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(6)]
again = hashlib.sha1(wav.cpu().numpy().tobytes()).hexdigest()[:16]
sw = {k: v for k, v in os.environ.items() if k.startswith("PB_")}
print(f"AB {sw} B={B} T={T} ms min={min(ts):.3f} all={['%.2f' % t for t in ts]} sha1={digest} stable={digest == again}", flush=True)
