"""Dev: time one decoder-shaped FFT block (B=32, L=2048) on the tensor-core path; PB_AT_ABLATE in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200.tte.fft import FFTBlock
blk = FFTBlock(256, 2, 1024, [9, 1], 0.1).eval().cuda(); blk.gemm_mode = 1
x = torch.randn(32, 2048, 256, device="cuda")
for _ in range(2): blk(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): blk(x)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("PB_AT_ABLATE", "0"), "block ms", e0.elapsed_time(e1) / 5)
