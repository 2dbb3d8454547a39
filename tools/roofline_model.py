"""Expected vs measured, per launch, for the vocoder forward (cfg2 by default) - CPU only, reads a committed ncu launch list.

The profiling recipe asks for the expected duration of every kernel BEFORE measuring; the gap is the optimisation signal.  This tool
states the two floors of every launch of one `CodeGenerator` forward and puts the launch list's measured duration beside them:

  hbm floor     bytes the launch has to move as implemented (split 16-bit activations = 4 B per element: conv1 reads x and writes xt;
                conv2 reads xt and the residual and writes; the ResBlock-closing conv2 also reads / writes the running sum; a fused pair
                reads x (with its halo) and writes once; weights once) / the copy roofline
  pipe floor    tcgen05 cycles of the launch from the measured cost of one MMA (profiles/r01_mma_probe.txt: M = 128, K = 16 costs
                max(N/2, 32 + N/4) cycles), two MMAs per k-step (x_hi against the stacked [w_hi; w_lo], x_lo against w_hi), spread over
                148 SMs at the SM clock

    python tools/roofline_model.py profiles/r01_launches_voc_cfg2.csv [--B 8 --T 1024] [--hbm 6650] [--mhz 1965]
    python tools/roofline_model.py profiles/r01_launches_tte_cfg3.csv --tte [--B 32 --S 256 --T 2048]

The launch order is the library's (voc_api.cu: embed, conv_pre, then per stage ups[i] and the ResBlocks k = 3, 7, 11, each (conv1, conv2) x 3
dilations - one fused launch per pair at C <= 32 - and conv_post).
"""
import argparse
import collections
import csv

C0, RATES, UPS_K, RB_K, N_DIL, SMS = 512, (5, 4, 4, 2, 2), (11, 8, 8, 4, 4), (3, 7, 11), 3, 148
EL = 4  # bytes per stored activation element (bf16 hi + bf16 lo)


def mma_cycles(n):
    return max(n / 2, 32 + n / 4)


def kstep_cycles(c_out):
    nt = min(c_out, 128)
    return (c_out // nt) * (mma_cycles(2 * nt) + mma_cycles(nt))      # per 128 rows, one tap, 16 input channels


def conv_cycles(rows, c_in, c_out, taps):
    return -(-rows // 128) * taps * (c_in // 16) * kstep_cycles(c_out)


def expected(B, T):
    """-> list of dict(name, bytes, cycles) in launch order."""
    out = [dict(name="embed", bytes=B * T * (16 + 256 * EL), cycles=0)]
    w = lambda cin, cout, k: cin * cout * k * 6                       # packed weights: stacked [hi; lo] + hi = 6 B per element
    out.append(dict(name="conv_pre 256->512 k7", bytes=B * T * (256 + 512) * EL + w(256, 512, 7), cycles=conv_cycles(B * T, 256, 512, 7)))
    L = T
    for i, (u, ku) in enumerate(zip(RATES, UPS_K)):
        cin, c = C0 >> i, C0 >> (i + 1)
        Lo = L * u
        out.append(dict(name=f"ups{i} {cin}->{c} k{ku} s{u}", bytes=B * (L * cin + Lo * c) * EL + w(cin, c, ku),
                        cycles=conv_cycles(B * Lo, cin, c, -(-ku // u))))
        rows = B * Lo
        act = rows * c * EL
        for j, k in enumerate(RB_K):
            for m in range(N_DIL):
                last = m == N_DIL - 1
                extra = (1 if (last and j > 0) else 0)              # running ResBlock sum read back
                if c == 16 and k >= 7:
                    # tc_pair16: super tile = 4 conv1 M tiles -> 4 (512 - (k - 1)) // 4 output rows; conv2 time-grouped (G = 4):
                    # G + k - 1 k-steps of {N = 128 stacked, N = 64}
                    m_out = 4 * ((512 - (k - 1)) // 4)
                    tiles = -(-rows // m_out)
                    cyc = tiles * (4 * k * kstep_cycles(16) + (4 + k - 1) * (mma_cycles(128) + mma_cycles(64)))
                    out.append(dict(name=f"s{i + 1} k{k} d{(1, 3, 5)[m]} pair C={c} tg", cycles=cyc,
                                    bytes=act * (1 + 1 + extra) + 2 * w(c, c, k)))
                elif c <= 32:                                         # fused pair: tiles of 128 xt rows -> 128 - (k - 1) output rows
                    tiles = -(-rows // (128 - (k - 1)))
                    cyc = tiles * 2 * k * (c // 16) * kstep_cycles(c)
                    halo = 128.0 / (128 - (k - 1))
                    out.append(dict(name=f"s{i + 1} k{k} d{(1, 3, 5)[m]} pair C={c}", cycles=cyc,
                                    bytes=act * (halo + 1 + extra) + 2 * w(c, c, k)))
                else:
                    out.append(dict(name=f"s{i + 1} k{k} d{(1, 3, 5)[m]} conv1 C={c}", bytes=2 * act + w(c, c, k),
                                    cycles=conv_cycles(rows, c, c, k)))
                    out.append(dict(name=f"s{i + 1} k{k} d1 conv2 C={c}", bytes=(3 + extra) * act + w(c, c, k),
                                    cycles=conv_cycles(rows, c, c, k)))
        L = Lo
    out.append(dict(name="conv_post 16->1 k7 + tanh", bytes=B * L * (16 * EL + 4), cycles=0))
    return out


def expected_tte(B, S, T, d=256, ff=1024, k1=9, heads=2, n_layers=4, n_codes=1000):
    """TTE forward (cfg3 by default) in launch order (tte_api.cu): embed, 4 encoder FFT blocks, speaker row, duration predictor,
    length regulator, 4 decoder FFT blocks, head, argmax.  GEMM inputs are split-fp16 rows (4 B per element), the residual
    stream is fp32; attention: Q.K^T is evaluated in both softmax passes, every product three times (split operands)."""
    F32 = 4
    out = [dict(name="embed + pe row", bytes=B * S * (8 + d * F32), cycles=0)]

    def gemm(name, M, K, taps, N, res):
        n_pad = -(-N // 128) * 128
        cyc = -(-M // 128) * (n_pad // 128) * taps * (K // 16) * kstep_cycles(128)
        byt = M * K * EL + M * N * F32 * (2 if res else 1) + K * taps * n_pad * 6
        return dict(name=name, bytes=byt, cycles=cyc)

    def block(tag, M, L):
        rows = []
        rows.append(dict(name=f"{tag} layernorm+split", bytes=M * d * (F32 + EL), cycles=0))
        rows.append(gemm(f"{tag} qkv {d}->{3 * d}", M, d, 1, 3 * d, False))
        att_flop = 9.0 * 2 * L * L * (d // heads) * heads * (M // L)              # (2 passes x 3 + 3) products
        rows.append(dict(name=f"{tag} attention L={L}", bytes=M * 3 * d * EL + M * d * EL, cycles=att_flop / 8192.0))
        rows.append(gemm(f"{tag} out-proj + res", M, d, 1, d, True))
        rows.append(dict(name=f"{tag} layernorm+split", bytes=M * d * (F32 + EL), cycles=0))
        rows.append(gemm(f"{tag} ffn k{k1} {d}->{ff} relu", M, d, k1, ff, False))
        rows.append(gemm(f"{tag} ffn k1 {ff}->{d} + res", M, ff, 1, d, True))
        return rows

    for i in range(n_layers):
        out += block(f"enc{i}", B * S, S)
    M = B * S
    out.append(dict(name="speaker row add", bytes=2 * M * d * F32, cycles=0))
    fp32_fma_per_clk = 128.0                                                     # per SM
    dp = dict(name="duration conv k3 (fp32 SIMT)", bytes=2 * M * d * F32, cycles=M * d * d * 3 / fp32_fma_per_clk * 1.0)
    out += [dp, dict(name="duration layernorm", bytes=2 * M * d * F32, cycles=0), dict(dp),
            dict(name="duration LN + head + round", bytes=M * d * F32, cycles=0), dict(name="duration scan", bytes=M * 16, cycles=0)]
    Md = B * T
    out.append(dict(name="length regulator + pe row", bytes=Md * d * F32 + M * d * F32, cycles=0))
    for i in range(n_layers):
        out += block(f"dec{i}", Md, T)
    out.append(dict(name="rows -> split", bytes=Md * d * (F32 + EL), cycles=0))
    head = gemm(f"head {d}->{n_codes} + argmax epilogue", Md, d, 1, n_codes, False)
    head["bytes"] = Md * d * EL + Md * 8          # reads the split rows, writes one 64-bit (value, index) key per row: no logits in HBM
    out.append(head)
    out.append(dict(name="argmax keys -> ids", bytes=Md * 16, cycles=0))
    return out


def read_launches(path, first="voc_embed"):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"name": r[ik]})
        d[r[im]] = float(r[iv].replace(",", ""))
    ids = list(per)
    start = max(i for i, k in enumerate(ids) if first in per[k]["name"])
    return [per[k] for k in ids[start:]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("launches")
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--hbm", type=float, default=6650.0, help="GB/s")
    ap.add_argument("--mhz", type=float, default=1965.0)
    ap.add_argument("--tte", action="store_true", help="TTE forward (defaults B=32 S=256 T=2048) instead of the vocoder")
    ap.add_argument("--S", type=int, default=256)
    a = ap.parse_args()
    if a.tte:
        B = 32 if a.B == 8 else a.B
        T = 2048 if a.T == 1024 else a.T
        exp = expected_tte(B, a.S, T)
    else:
        exp = expected(a.B, a.T)
    got = [d for d in read_launches(a.launches, "embed_pos" if a.tte else "voc_embed") if "at::" not in d["name"]]
    assert len(got) >= len(exp), (len(got), len(exp))
    print(f"{'launch':34s} {'meas us':>8s} {'hbm us':>7s} {'pipe us':>8s} {'bound':>6s} {'floor/meas':>10s} {'dram MB':>8s} {'model MB':>8s}")
    stage = collections.OrderedDict()
    for e, g in zip(exp, got):
        us = g["gpu__time_duration.sum"] / 1e3
        t_h = e["bytes"] / (a.hbm * 1e3)
        t_t = e["cycles"] / SMS / a.mhz
        floor = max(t_h, t_t)
        dram = (g.get("dram__bytes_read.sum", 0) + g.get("dram__bytes_write.sum", 0)) / 1e6
        print(f"{e['name']:34s} {us:8.1f} {t_h:7.1f} {t_t:8.1f} {'hbm' if t_h >= t_t else 'pipe':>6s} {floor / us:10.2f} {dram:8.1f} {e['bytes'] / 1e6:8.1f}")
        key = e["name"].split()[0]
        if a.tte:
            key = key[:3] if key[:3] in ("enc", "dec") else "other"
            if key != "other":
                key += " attention" if "attention" in e["name"] else (" gemm" if ("->" in e["name"] or "proj" in e["name"]) else " layernorm")
        else:
            key = key if key[0] == "s" and key[1:].isdigit() else "other"
        s = stage.setdefault(key, [0.0, 0.0, 0.0, 0.0])
        s[0] += us; s[1] += t_h; s[2] += t_t; s[3] += floor
    print()
    print(f"{'section':14s} {'meas us':>9s} {'hbm us':>9s} {'pipe us':>9s} {'floor us':>9s} {'floor/meas':>10s}")
    for k, s in list(stage.items()) + [("total", [sum(v[i] for v in stage.values()) for i in range(4)])]:
        print(f"{k:14s} {s[0]:9.1f} {s[1]:9.1f} {s[2]:9.1f} {s[3]:9.1f} {s[3] / s[0]:10.2f}")


if __name__ == "__main__":
    main()
