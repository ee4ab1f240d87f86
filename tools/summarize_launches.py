#!/usr/bin/env python
"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --frames 1`: per-kernel launch count,
total device time and share for the LAST frame processed (the list is cut at the kWalk launches: one per frame).
Usage: summarize_launches.py launches.csv out.csv [comment]"""
import collections
import csv
import json
import sys

rows, hdr = [], None
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
        name = d["Kernel Name"].split("(")[0]
        for junk in ("pccb200::", "(anonymous namespace)::", "<unnamed>::", "void "):
            name = name.replace(junk, "")
        rows.append((name, v))
walks = [i for i, (k, _) in enumerate(rows) if k.startswith("kWalk")]
# a frame's launches: from after the previous frame's last launch ... the list has no frame marker, so split half-way between walks
if len(walks) >= 2:
    # launches before a walk (a1-a4 + orient prepare) belong to the same frame as the walk: cut where the previous frame ended
    tail = len(rows) - walks[-1]            # launches from the last walk to the end (post-walk stages of the last frame)
    start = walks[-2] + tail                # the previous frame had the same number of post-walk launches (same data)
else:
    start = 0
sel = rows[start:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in sel:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("# %s\n" % (sys.argv[3] if len(sys.argv) > 3 else "ncu launch list"))
    f.write("# gpu__time_duration.sum per kernel; cold-cache, serialised: compare SHARES, not absolutes. launches %d, total %.1f ms, non-walk %.2f ms\n"
            % (len(sel), tot, tot - sum(v[1] for k, v in agg.items() if k.startswith("kWalk"))))
    f.write("kernel,launches,total_ms,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.4f\n" % (k[:80], v[0], v[1], v[1] / tot))
json.dump({"launches_per_frame": len(sel)}, sys.stdout)
print()
