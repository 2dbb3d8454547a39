"""Dev: tensor-core conv1d / conv_transpose1d vs torch on the GPU.  usage: dev_tc_conv.py bo_mode [quick]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from parrot_tts_b200 import ops

bo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = [(16, 3, 1, 300), (32, 3, 1, 300), (64, 3, 1, 300), (64, 1, 1, 200), (16, 7, 3, 1000), (32, 11, 5, 777), (64, 7, 5, 500),
         (128, 3, 3, 260), (256, 7, 1, 200), (256, 11, 5, 130), (16, 11, 5, 5000)]
if len(sys.argv) > 2:
    cases = cases[:int(sys.argv[2])]
dev = "cuda:0"
for (C, K, D, L) in cases:
    g = torch.Generator().manual_seed(C * 100 + K * 10 + D)
    B = 2
    x = torch.randn(B, C, L, generator=g)
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g)
    r = torch.randn(B, C, L, generator=g)
    ref = F.leaky_relu(F.conv1d(F.leaky_relu(x, 0.1), w, b, padding=(K - 1) * D // 2, dilation=D) + r, 0.1)
    try:
        got = ops.conv1d_tc(x.to(dev), w.to(dev), b.to(dev), dilation=D, res=r.to(dev), pre_slope=0.1, post_slope=0.1, single_pass=bool(bo)).cpu()
        err = (got - ref).abs()
        print(f"conv C={C} K={K} D={D} L={L} bo={bo}: max err {float(err.max()):.3e} mean {float(err.mean()):.3e} (ref max {float(ref.abs().max()):.2f})", flush=True)
    except Exception as e:
        print(f"conv C={C} K={K} D={D} L={L} bo={bo}: EXC {e}", flush=True)
        break
for (Cin, Cout, K, U, L) in [(32, 16, 4, 2, 300), (64, 32, 4, 2, 300), (128, 64, 8, 4, 200), (512, 256, 11, 5, 150)]:
    g = torch.Generator().manual_seed(Cin + K)
    x = torch.randn(2, Cin, L, generator=g)
    w = torch.randn(Cin, Cout, K, generator=g) / (Cin * K / U) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=U, padding=(K - U) // 2)
    try:
        got = ops.conv_transpose1d_tc(x.to(dev), w.to(dev), b.to(dev), stride=U, pre_slope=0.1, single_pass=bool(bo)).cpu()
        err = (got - ref).abs()
        print(f"convtr Cin={Cin} Cout={Cout} K={K} U={U} bo={bo}: max err {float(err.max()):.3e} mean {float(err.mean()):.3e}", flush=True)
    except Exception as e:
        print(f"convtr Cin={Cin} K={K} U={U} bo={bo}: EXC {e}", flush=True)
        break
