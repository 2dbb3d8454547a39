"""ORACLE-side experiment (test infrastructure, not product code; CPU only): what would cheaper operand formats cost in waveform error?

The vocoder convolutions run today as 3 tensor passes per product (split bf16: x_hi.w_hi + x_lo.w_hi + x_hi.w_lo, ~2^-16 per product).
This script replays the oracle graph with the operands of every
convolution rounded the way a cheaper tcgen05 scheme would see them, and reports max |wav - wav_fp32| against the 1e-3 abs bar:

  tf32_trunc   kind::tf32, activations fed as raw fp32 (the MMA ignores the low 13 mantissa bits), weights rounded to tf32 (RNE)
  tf32_rne     kind::tf32, both operands rounded to nearest tf32
  x22_w11      2 passes of kind::f16: activation as fp16 hi + lo (22 bits), weight as ONE fp16 (11 bits, RNE)
  x16_w16      today's 3-pass split bf16 (x and w to 16 mantissa bits, the lo.lo term dropped)
  bf16         1 pass plain bf16

    python oracle/precision_emulation.py [--units 48] [--batch 2] [--gain 1.0 6.0]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import voc_oracle  # noqa: E402
from parrot_tts_b200 import configs, synth  # noqa: E402


def keep_bits(x, bits, rne=True):
    """Keep ``bits`` explicit mantissa bits of fp32 (tf32: 10, bf16: 7)."""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    if rne:
        half = (1 << (drop - 1)) - 1
        i = i + half + ((i >> drop) & 1)
    return (i & ~((1 << drop) - 1)).view(torch.float32)


def split(x, bits_hi, bits_lo):
    hi = keep_bits(x, bits_hi)
    return hi + keep_bits(x - hi, bits_lo)


SCHEMES = {
    "tf32_trunc": (lambda x: keep_bits(x, 10, rne=False), lambda w: keep_bits(w, 10)),
    "tf32_rne": (lambda x: keep_bits(x, 10), lambda w: keep_bits(w, 10)),
    "x22_w11": (lambda x: split(x, 10, 10), lambda w: keep_bits(w, 10)),
    "x16_w16": (lambda x: split(x, 7, 7), lambda w: split(w, 7, 7)),
    "bf16": (lambda x: keep_bits(x, 7), lambda w: keep_bits(w, 7)),
}


def run(scheme, sd, h, code, spkr, min_channels=0):
    """``min_channels``: the cheap scheme only where the convolution has at least that many input channels (the wide,
    tensor-bound stages); everything else keeps today's x16_w16."""
    if scheme is None:
        return voc_oracle.code_generator_forward(sd, h, code, spkr)
    c1, ct = F.conv1d, F.conv_transpose1d

    def pick(x):
        return SCHEMES[scheme] if x.shape[1] >= min_channels else SCHEMES["x16_w16"]

    F.conv1d = lambda x, w, b=None, **kw: c1(pick(x)[0](x), pick(x)[1](w), b, **kw)
    F.conv_transpose1d = lambda x, w, b=None, **kw: ct(pick(x)[0](x), pick(x)[1](w), b, **kw)
    try:
        return voc_oracle.code_generator_forward(sd, h, code, spkr)
    finally:
        F.conv1d, F.conv_transpose1d = c1, ct


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=48)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--gain", type=float, nargs="+", default=[1.0, 6.0])
    ap.add_argument("--seeds", type=int, default=2)
    a = ap.parse_args()
    h = configs.default_vocoder_config()
    for gain in a.gain:
        for seed in range(a.seeds):
            sd = synth.synth_vocoder_state_dict(h, seed=seed, out_gain=gain)
            g = torch.Generator().manual_seed(seed)
            code = torch.randint(0, 1000, (a.batch, a.units), generator=g)
            spkr = torch.randint(0, 10, (a.batch, 1), generator=g)
            ref = run(None, sd, h, code, spkr)
            row = [f"gain {gain:g} seed {seed} |y|max {float(ref.abs().max()):.3f}"]
            for name in SCHEMES:
                err = float((run(name, sd, h, code, spkr) - ref).abs().max())
                row.append(f"{name} {err:.2e}")
            for name, cmin in (("x22_w11", 256), ("x22_w11", 128), ("tf32_rne", 256), ("tf32_rne", 128)):
                err = float((run(name, sd, h, code, spkr, cmin) - ref).abs().max())
                row.append(f"{name}@C>={cmin} {err:.2e}")
            print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
