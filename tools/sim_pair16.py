"""CPU only: discrete-event model of tc_pair16_kernel's pipeline (parrot_tts_b200/csrc/tc_pair.cu) - an in-order tensor pipe fed by the
conv1 and conv2 issuing warps, four epilogue groups that run epilogue A of super tile i + 1 before epilogue B of tile i, two stages of
every buffer.  It answers "what period per super tile would this dependency structure give if the hardware behaved as assumed":
    python tools/sim_pair16.py            # k = 11 at C = 16: tensor work 4 x 869 + 1568 cycles per super tile
The model says tensor-bound (~5 070 cycles); the launch measures 8 750 - i.e. the dependency structure is not
what costs the difference."""
import argparse


def run(T1m=869, T2=1568, hop=300, a=300, b=400, wait=150, commit=150, ntiles=40, qdepth=None):
    """T1m / T2: tensor cycles of one conv1 M tile / of conv2; hop: barrier completion -> waiter running; a / b: busy time of epilogue
    A / B; wait / commit: the issuing thread's cost of one mbarrier wait / one tcgen05.commit; qdepth: cycles of MMA work the pipe
    accepts ahead (None = unbounded queue)."""
    pipe_free = 0

    def pipe(t_issue, dur):
        nonlocal pipe_free
        start = max(pipe_free, t_issue)
        pipe_free = start + dur
        return pipe_free

    c1_t = c2_t = 0
    grp_t = [0] * 4
    a_done, c1_done, c2_done = {}, {}, {}
    c1_pos, c2_pos, g_pos = (0, 0), 0, [("A", 0)] * 4
    while True:
        cands = []
        i, m = c1_pos
        if i < ntiles:
            need = a_done.get((i - 2, m)) if i >= 2 else 0          # conv1(i, m) overwrites the accumulator epilogue A(i - 2, m) read
            if need is not None:
                cands.append((max(c1_t, need + (hop if i >= 2 else 0)) + 2 * wait, "c1"))
        if c2_pos < ntiles:
            ns = [a_done.get((c2_pos, g)) for g in range(4)]        # conv2(i) needs every group's xt rows
            if all(n is not None for n in ns):
                cands.append((max(c2_t, max(ns) + hop) + wait, "c2"))
        for g in range(4):
            kind, i = g_pos[g]
            if i >= ntiles:
                continue
            need = c1_done.get((i, g)) if kind == "A" else c2_done.get(i)
            if need is not None:
                cands.append((max(grp_t[g], need + hop), f"g{g}"))
        if not cands:
            break
        t, who = min(cands)
        if who == "c1":
            i, m = c1_pos
            fin = pipe(t, T1m)
            c1_done[(i, m)] = fin
            c1_t = t + commit if qdepth is None else max(t + commit, fin - qdepth)
            c1_pos = (i, m + 1) if m < 3 else (i + 1, 0)
        elif who == "c2":
            fin = pipe(t, T2)
            c2_done[c2_pos] = fin
            c2_t = t + commit if qdepth is None else max(t + commit, fin - qdepth)
            c2_pos += 1
        else:
            g = int(who[1])
            kind, i = g_pos[g]
            if kind == "A":
                a_done[(i, g)] = grp_t[g] = t + a
                g_pos[g] = ("A", 1) if i == 0 else ("B", i - 1)      # A(0), then A(i + 1) before B(i)
            else:
                grp_t[g] = t + b
                g_pos[g] = ("A", i + 2) if i + 2 < ntiles else (("B", i + 1) if i + 1 < ntiles else ("A", ntiles))
    return max(grp_t) / ntiles


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=11)
    a = ap.parse_args()
    t1m, t2 = a.k * 79, (4 + a.k - 1) * 112
    print(f"k = {a.k}: conv1 {t1m} cycles per M tile (x4), conv2 {t2}: tensor work {4 * t1m + t2} per super tile")
    for kw in (dict(), dict(hop=600), dict(a=800, b=1200), dict(qdepth=200), dict(T1m=0, T2=0)):
        args = dict(T1m=t1m, T2=t2)
        args.update(kw)
        print(f"  {kw or 'defaults'}: {run(**args):.0f} cycles per super tile")
