"""Per-launch durations of one vocoder forward under several PB_TC_ABLATE settings, side by side.
usage: ablate_table.py gpurun_out/abl_<n>.csv ...   (files from tools/gpu_ablate.sh; ncu launch lists of tools/prof_voc.py 8 1024 1 0)"""
import csv, sys, collections, re, os

def load(f):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        per[r[iid]] = (r[ik].split("(")[0].replace("void ", "").replace("pb::", ""), float(r[iv].replace(",", "")) / 1e3)
    out = [v for v in per.values() if v[0].startswith(("tc_conv_kernel", "tc_pair_kernel", "tc_pair16_kernel", "conv_post", "voc_embed"))]
    return out

def label(i):      # launch index inside one tensor-core forward (80 launches: embed, conv_pre, 5 x (ups + 18 or 9 resblock launches), conv_post)
    names = ["embed", "conv_pre"]
    for s, n in enumerate((18, 18, 18, 9, 9)):
        names.append(f"ups{s}")
        ks = (3, 7, 11)
        if n == 18:
            for k in ks:
                for d in (1, 3, 5):
                    names += [f"s{s+1} k{k} d{d} c1", f"s{s+1} k{k} d{d} c2"]
        else:
            for k in ks:
                for d in (1, 3, 5):
                    names.append(f"s{s+1} k{k} d{d} pair")
    names.append("conv_post")
    return names[i] if i < len(names) else str(i)

files = sys.argv[1:]
cols = [load(f) for f in files]
tags = [re.sub(r".*abl_(\w+)\.csv", r"\1", f) for f in files]
n = min(len(c) for c in cols)
print(f"{'launch':22s} " + " ".join(f"{t:>8s}" for t in tags))
tot = [0.0] * len(cols)
for i in range(n):
    print(f"{label(i):22s} " + " ".join(f"{c[i][1]:8.1f}" for c in cols))
    for j, c in enumerate(cols):
        tot[j] += c[i][1]
print(f"{'total us':22s} " + " ".join(f"{t:8.0f}" for t in tot))
