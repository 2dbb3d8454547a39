import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from parrot_tts_b200 import ops
DEV = "cuda:0"
y = torch.rand(1, 1, 4096) * 2 - 1
p = ops.wav_to_pcm16(y.to(DEV)); torch.cuda.synchronize(); print("pcm ok", flush=True)
for (C, L, K, D) in [(64, 333, 3, 1), (16, 1531, 3, 1), (32, 700, 7, 3)]:
    x = torch.randn(2, C, L); w = torch.randn(C, C, K) / (C * K) ** 0.5; b = torch.randn(C)
    ref = F.conv1d(x, w, b, padding=(K - 1) * D // 2, dilation=D)
    got = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), dilation=D)
    torch.cuda.synchronize()
    print("conv", C, L, K, D, "maxerr", float((got.cpu() - ref).abs().max()), flush=True)
