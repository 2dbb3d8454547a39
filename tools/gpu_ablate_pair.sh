#!/bin/bash
# per-launch durations of the 18 fused-pair launches (stages 4-5) of one 8 x 1024-unit vocoder forward under PB_TP_ABLATE settings
# (dev switch of tc_pair.cu: 1 no MMAs, 2 no output stores, 4 no residual / running-sum loads, 8 no x loads, 16 no xt stores,
# 32 no epilogue arithmetic)
mkdir -p gpurun_out
for a in "$@"; do
  PB_TP_ABLATE=$a timeout 300 ncu -k regex:tc_pair --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ablp_$a.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "ablate $a rc=$?"
done
python - "$@" <<'PY' > gpurun_out/ablate_pair_table.txt
import csv, sys
names = [f"s{s} k{k} d{d}" for s in (4, 5) for k in (3, 7, 11) for d in (1, 3, 5)]
cols = []
for a in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/ablp_{a}.csv") if l.startswith('"'))]
    h = rows[0]
    cols.append([float(r[h.index("Metric Value")].replace(",", "")) / 1e3 for r in rows[1:] if r[h.index("Metric Name")] == "gpu__time_duration.sum"])
print(f"{'launch':12s} " + " ".join(f"{a:>7s}" for a in sys.argv[1:]))
for i, n in enumerate(names):
    print(f"{n:12s} " + " ".join(f"{c[i]:7.1f}" if i < len(c) else "    nan" for c in cols))
print(f"{'total us':12s} " + " ".join(f"{sum(c[:18]):7.0f}" for c in cols))
PY
cat gpurun_out/ablate_pair_table.txt
