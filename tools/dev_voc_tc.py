"""Dev: CodeGenerator on the tensor-core path vs the CPU oracle, then timing at B=8 x T=1024."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import voc_oracle
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator

dev = "cuda:0"
h = configs.default_vocoder_config()
sd = synth.synth_vocoder_state_dict(h, seed=11)
g = torch.Generator().manual_seed(5)
code = torch.randint(0, 1000, (2, 50), generator=g)
spkr = torch.randint(0, 10, (2, 1), generator=g)
ref = voc_oracle.code_generator_forward(sd, h, code, spkr)
for mode in (0, 1, 2):
    gen = CodeGenerator(h); gen.conv_mode = mode
    gen.load_state_dict(sd); gen = gen.to(dev)
    got = gen(code=code.to(dev), spkr=spkr.to(dev)).cpu()
    e = (got - ref).abs()
    print(f"mode {mode}: max err {float(e.max()):.3e} mean {float(e.mean()):.3e} ref absmax {float(ref.abs().max()):.3f} launches {gen.last_launches}", flush=True)
    lens = torch.tensor([50, 23])
    got = gen(code=code.to(dev), spkr=spkr.to(dev), lens=lens).cpu()
    r1 = voc_oracle.code_generator_forward(sd, h, code[1:2, :23], spkr[1:2])
    print(f"   ragged item1 err {float((got[1:2, :, :320*23] - r1).abs().max()):.3e} tail absmax {float(got[1, :, 320*23:].abs().max()):.1e}", flush=True)
B, T = 8, 1024
code = torch.randint(0, 1000, (B, T), device=dev); spkr = torch.randint(0, 10, (B, 1), device=dev)
for mode in (1, 2):
    gen = CodeGenerator(h); gen.conv_mode = mode
    gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); gen = gen.to(dev)
    for _ in range(3): wav = gen(code=code, spkr=spkr)
    torch.cuda.synchronize()
    gen.set_timing(True, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for i in range(5):
        ev[i].record(); wav = gen(code=code, spkr=spkr)
    ev[5].record(); torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    ms = min(ts)
    print(f"mode {mode}: B={B} T={T} ms/forward={ms:.3f} samples/s={B*T*320/ms*1e3:.4e} eff TFLOP/s={321656832.0*B*T/ms/1e9:.1f} HBM-alg GB/s={3616000.0*B*T/ms/1e6:.0f}")
    print("   sections:", {k: round(v, 3) for k, v in gen.section_ms().items()}, flush=True)
