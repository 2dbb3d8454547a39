"""Per-kernel totals from an ncu launch list CSV (gpu__time_duration + dram bytes). usage: launch_table.py file [--each]"""
import csv, sys, collections
f = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[ik]})
    d[r[im]] = float(r[iv].replace(",", ""))
# --last NAME: keep only the launches from the last kernel whose name contains NAME (= the last forward pass)
if "--last" in sys.argv:
    key = sys.argv[sys.argv.index("--last") + 1]
    ids = list(per)
    start = max(i for i, k in enumerate(ids) if key in per[k]["name"])
    per = collections.OrderedDict((k, per[k]) for k in ids[start:])
tot = collections.OrderedDict()
for i, d in per.items():
    n = d["name"].split("(")[0].replace("void ", "").replace("pb::", "")
    t = tot.setdefault(n, [0, 0.0, 0.0])
    t[0] += 1; t[1] += d.get("gpu__time_duration.sum", 0) / 1e3; t[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    if "--each" in sys.argv:
        print(f"{i:>4} {n[:44]:44s} {d.get('gpu__time_duration.sum', 0)/1e3:9.1f} us  dram {((d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0))/1e6):9.1f} MB")
all_t = sum(t[1] for t in tot.values())
print(f"{'kernel':44s} {'launches':>8s} {'us':>10s} {'share':>7s} {'dram MB':>10s}")
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:44]:44s} {t[0]:8d} {t[1]:10.1f} {100*t[1]/all_t:6.1f}% {t[2]/1e6:10.1f}")
print(f"{'total':44s} {sum(t[0] for t in tot.values()):8d} {all_t:10.1f}")
