#!/bin/bash
# scratch: coarse variant (one x tile / one wait / one commit per super tile) vs the product, time-grouped kernel at every k
mkdir -p gpurun_out
for v in "" coarse; do PB_LIB_VARIANT=$v PB_TP_TOEPLITZ=1 python tools/dev_ab.py 8 1024; PB_LIB_VARIANT=$v PB_TP_TOEPLITZ=1 python tools/dev_ab.py 3 37; done 2>&1 | grep AB
run() {  # name env...
  local name=$1; shift
  env "$@" timeout 300 ncu -k regex:tc_pair --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ablp_$name.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$name rc=$?"
}
run t0 PB_TP_TOEPLITZ=0
run t1 PB_TP_TOEPLITZ=1
run c1 PB_TP_TOEPLITZ=1 PB_LIB_VARIANT=coarse
run c1a111 PB_TP_TOEPLITZ=1 PB_LIB_VARIANT=coarse PB_TP_ABLATE=111
python - t0 t1 c1 c1a111 <<'PY'
import csv, sys
names = [f"s{s} k{k} d{d}" for s in (4, 5) for k in (3, 7, 11) for d in (1, 3, 5)]
cols = []
for a in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/ablp_{a}.csv") if l.startswith('"'))]
    h = rows[0]
    cols.append([float(r[h.index("Metric Value")].replace(",", "")) / 1e3 for r in rows[1:] if r[h.index("Metric Name")] == "gpu__time_duration.sum"])
print(f"{'launch':12s} " + " ".join(f"{a:>7s}" for a in sys.argv[1:]))
for i, n in enumerate(names):
    if i >= 9: print(f"{n:12s} " + " ".join(f"{c[i]:7.1f}" if i < len(c) else "    nan" for c in cols))
PY
