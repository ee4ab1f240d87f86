#!/usr/bin/env python
"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`) of
`bench.py --frames 1`: per-kernel launch count, total device time, share and (when captured) DRAM bytes for the LAST-BUT-ONE
frame processed (the list is cut at the kWalk launches, one per frame; the last frame of a bench run is the size probe, which
stops early). Usage: summarize_launches.py launches.csv out.csv [comment]"""
import collections
import csv
import json
import sys

TIME = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
by, hdr = collections.OrderedDict(), None
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = d["Kernel Name"].split("(")[0]
        for junk in ("pccb200::", "(anonymous namespace)::", "<unnamed>::", "void "):
            name = name.replace(junk, "")
        e = by.setdefault(d["ID"], {"name": name, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        m, u = d["Metric Name"], d["Metric Unit"]
        if m == "gpu__time_duration.sum":
            e["ms"] = v * TIME.get(u, 1e-6)
        elif m == "dram__bytes_read.sum":
            e["rd"] = v * BYTES.get(u, 1)
        elif m == "dram__bytes_write.sum":
            e["wr"] = v * BYTES.get(u, 1)
rows = list(by.values())
walks = [i for i, e in enumerate(rows) if e["name"].startswith("kWalk")]
# a frame's launches: the ones before its walk (a1-a4 + orient prepare) and after it up to the next frame's first launch; all
# frames of the run are the same cloud, so frame k spans [walk_k - pre, walk_{k+1} - pre) with pre = launches before the first walk
if len(walks) >= 2:
    pre = walks[0]
    sel = rows[walks[-2] - pre: walks[-1] - pre]
else:
    sel = rows
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for e in sel:
    a = agg[e["name"]]
    a[0] += 1
    a[1] += e["ms"]
    a[2] += e["rd"]
    a[3] += e["wr"]
tot = sum(v[1] for v in agg.values())
walk = sum(v[1] for k, v in agg.items() if k.startswith("kWalk"))
with open(sys.argv[2], "w") as f:
    f.write("# %s\n" % (sys.argv[3] if len(sys.argv) > 3 else "ncu launch list"))
    f.write("# gpu__time_duration.sum per kernel; cold-cache, serialised: compare SHARES, not absolutes. launches %d, total %.1f ms, non-walk %.2f ms; "
            "dram bytes = dram__bytes_read.sum / dram__bytes_write.sum summed over the kernel's launches (writes absorbed by the 126 MB L2 show as 0)\n"
            % (len(sel), tot, tot - walk))
    f.write("kernel,launches,total_ms,share,dram_read_MB,dram_write_MB,dram_GBps\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.4f,%.1f,%.1f,%.0f\n" % (k[:80], v[0], v[1], v[1] / tot, v[2] / 1e6, v[3] / 1e6, (v[2] + v[3]) / 1e9 / (v[1] / 1e3) if v[1] else 0))
json.dump({"launches_per_frame": len(sel), "non_walk_ms": round(tot - walk, 2)}, sys.stdout)
print()
