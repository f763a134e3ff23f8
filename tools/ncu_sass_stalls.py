"""Developer tool: top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output.
    python tools/ncu_sass_stalls.py file.csv [N]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # the export repeats the kernel once per view: keep the first listing
    if len(r) >= len(hdr) - 1:
        data.append(r + [""] * (len(hdr) - len(r)))
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = sum(float(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, " instructions", len(data))
agg = {c: sum(float(r[ix[c]] or 0) for r in data) for c in stall_cols}
print("by reason:", {k.replace("stall_", ""): round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v / tot > 0.01})
order = sorted(range(len(data)), key=lambda i: -float(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = float(r[ix["# Samples"]] or 0)
    top = sorted(((float(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (i, 100 * s / tot, r[ix["Source"]][:70], " ".join("%s=%d" % (c.replace("stall_", ""), v) for v, c in top if v)))
