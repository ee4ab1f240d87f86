"""Tiny driver for ncu captures: one segment_frame call (a1-a11) on a synthetic figure. Usage: python profiles/run_one_frame.py [scale]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bindings  # noqa: E402
import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
xyz, rgb = synth.figure(scale=scale, seed=0, frame=0)
p = bindings.Product(0)
prm = bindings.ctc_seg_params(bits=10, iterations=10, weight=p.weight_normal(xyz, 11))
out = p.segment_frame(xyz, rgb, prm)
print(len(xyz), "points,", len(out["patches"].patches), "patches")
