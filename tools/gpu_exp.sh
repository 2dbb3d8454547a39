#!/bin/bash
# scratch: residual rows from the x tile in shared memory (tap-by-tap fused kernel) - correctness, then per launch, same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_switches_gpu.py tests/test_voc_gpu.py -x -q -m gpu -k "residual_from or time_grouped or golden or ragged or rectangle or other_" 2>&1 | tail -4
run() {  # name env...
  local name=$1; shift
  env "$@" timeout 300 ncu -k regex:tc_pair --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ablp_$name.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$name rc=$?"
}
run r0 PB_TP_RES_SMEM=0
run r1 PB_TP_RES_SMEM=1
run r0b PB_TP_RES_SMEM=0
python - r0 r1 r0b <<'PY'
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
print(f"{'18 pairs':12s} " + " ".join(f"{sum(c[:18]):7.0f}" for c in cols))
PY
