"""Top stall-sample instructions from an ncu source-page CSV. usage: ncu_hot.py file [n]"""
import csv, sys
f = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(f)))
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try: s = int(r[isamp])
    except: continue
    data.append((idx, s, r))
tot = sum(s for _, s, _ in data)
print("total samples", tot, "instructions", len(data))
for idx, s, r in sorted(data, key=lambda x: -x[1])[:n]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {s:6d} {100*s/tot:5.1f}% exe={r[iexe]:>8} {r[isrc].strip()[:70]:70s} {st}")
