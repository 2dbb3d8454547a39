"""CPU check of the PB_VOC_CHAINS=1 schedule (voc_api.cu, voc_forward_tc): a mirror of its launch / event sequence, with every
buffer access, and a proof by transitive closure that each conflicting pair of accesses (RAW / WAR / WAW on a workspace slot) is ordered
by stream order or an event, and that every side-stream launch precedes conv_post (all forks joined: capturable as a CUDA graph).
Keep in step with voc_forward_tc when the schedule changes.      python tools/check_chain_schedule.py
"""
import itertools

NK, ND, NUPS, FUSED_FROM = 3, 3, 5, 3          # kernel sizes, dilations, stages; stages >= FUSED_FROM use one fused launch per pair


def schedule():
    n_slots = 4 + 3 * NK
    ops, edges, last_on, events, pending = [], [], {}, {}, {}

    def launch(stream, reads, writes, name):
        i = len(ops)
        ops.append((stream, set(reads), set(writes), name))
        if stream in last_on:
            edges.append((last_on[stream], i))
        for w in pending.pop(stream, []):
            edges.append((w, i))
        last_on[stream] = i

    def record(ev, stream):
        events[ev] = last_on.get(stream)

    def wait(stream, ev):
        if events.get(ev) is not None:
            pending.setdefault(stream, []).append(events[ev])

    slot, in_slot = list(range(n_slots)), -1
    launch("st", ["code"], ["x0"], "embed")
    launch("st", ["x0"], ["h0"], "conv_pre")
    cur_in = "h0"
    for i in range(NUPS):
        buf = [s for s in slot if s != in_slot][:n_slots - 1]
        X, XS, Y = buf[0], buf[4], buf[5]
        launch("st", [cur_in], [X], f"ups{i}")
        record("fork", "st")
        for j in range(NK - 1):
            wait(f"chain{j}", "fork")
        for j in range(NK):
            XT, XA, XB = (buf[1], buf[2], buf[3]) if j == 0 else (buf[6 + 3 * (j - 1)], buf[7 + 3 * (j - 1)], buf[8 + 3 * (j - 1)])
            sj = f"chain{j}" if j + 1 < NK else "st"
            cur = X
            for m in range(ND):
                fused = i >= FUSED_FROM
                if not fused:
                    launch(sj, [cur], [XT], f"s{i} k{j} m{m} conv1")
                src = [cur] if fused else [XT, cur]
                if m + 1 < ND:
                    dst = XB if cur == XA else XA
                    launch(sj, src, [dst], f"s{i} k{j} m{m} conv2")
                    cur = dst
                else:
                    if j > 0:
                        wait(sj, f"last{j - 1}")
                    launch(sj, src + ([XS] if j > 0 else []), [XS if j + 1 < NK else Y], f"s{i} k{j} closing conv2")
                    if j + 1 < NK:
                        record(f"last{j}", sj)
        cur_in, in_slot = Y, Y
    launch("st", [cur_in], ["wav"], "conv_post")
    return ops, edges


def main():
    ops, edges = schedule()
    n = len(ops)
    succ = [[] for _ in range(n)]
    for a, b in edges:
        succ[a].append(b)
    reach = [set() for _ in range(n)]
    for a in range(n - 1, -1, -1):
        for b in succ[a]:
            reach[a].add(b)
            reach[a] |= reach[b]
    bad = 0
    for a, b in itertools.combinations(range(n), 2):
        _, ra, wa, na = ops[a]
        _, rb, wb, nb = ops[b]
        if ((wa & (rb | wb)) or (ra & wb)) and b not in reach[a]:
            bad += 1
            print("UNORDERED:", na, "|", nb)
    joined = all((n - 1) in reach[a] for a in range(n - 1))
    print(f"{n} launches, {bad} unordered conflicting pairs, every launch precedes conv_post: {joined}")
    raise SystemExit(0 if bad == 0 and joined else 1)


if __name__ == "__main__":
    main()
