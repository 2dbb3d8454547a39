"""DRAM traffic of the ResBlock conv family from ncu launch lists (tools/gpu_prof_r2.sh): per workload the average
dram__bytes_read + dram__bytes_write over the 72 ResBlock launches of ONE CodeGenerator forward -> profiles/traffic.json
(bench.py reports it as roofline.traffic beside the layer-wise algorithmic bytes; a profiler pass, never a timed one).
usage: ncu_traffic.py cfg2=profiles/r02_launches_voc_cfg2.csv cfg3=profiles/r02_launches_voc_cfg3.csv"""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(f):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"name": r[ik].split("(")[0].replace("void ", "").replace("pb::", "")})
        d[r[im]] = float(r[iv].replace(",", ""))
    return [d for d in per.values() if d["name"].startswith(("tc_conv_kernel", "tc_pair_kernel", "tc_pair16_kernel", "conv_post", "voc_embed"))]


def resblock_indices():      # inside one tensor-core forward: embed, conv_pre, 5 x (ups, 18 or 9 resblock launches), conv_post
    idx, i = [], 2
    for n in (18, 18, 18, 9, 9):
        i += 1                 # ups
        idx += list(range(i, i + n))
        i += n
    return idx


out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over tools/prof_voc.py (one forward); "
                 "average over the 72 ResBlock launches; files under profiles/"}
for arg in sys.argv[1:]:
    name, f = arg.split("=")
    L = launches(f)[-80:]
    assert len(L) == 80, (f, len(L))
    rb = [L[i] for i in resblock_indices()]
    assert len(rb) == 72 and all(d["name"].startswith(("tc_conv", "tc_pair")) for d in rb)
    tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in rb)
    us = sum(d["gpu__time_duration.sum"] for d in rb) / 1e3
    out[name] = tot / 72
    out[name + "_detail"] = {"file": os.path.relpath(f, ROOT), "resblock_dram_bytes_per_forward": tot, "resblock_us_per_forward_under_ncu": us,
                             "forward_us_under_ncu": sum(d["gpu__time_duration.sum"] for d in L) / 1e3}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
