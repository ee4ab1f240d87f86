"""Generates tests/golden/gof_fullsize.json from the REFERENCE ITSELF (oracle/_ref/libtmc2ref.so) at the sizes bench.py runs:
  ai_r3    2 frames of figure(scale=0.626) (0.83 Mpts, the default bench workload), CTC all-intra r3, I=50
  ra_r5    a 4-frame random-access GOF at 0.83 Mpts (occupancyPrecision 2, global patch allocation), I=50
  vox11    one 11-bit frame of ~2.9 Mpts (2560-wide canvas), I=20 (basketball_player cfg)
Run in the build container:  python tests/golden/make_golden_fullsize.py   (about five minutes, single thread)
Only sha256 digests of every hot-path product are committed; the GPU tests (tests/test_gpu_fullsize.py) and bench.py's
in-run parity check compare the CUDA path's products with them, so the benchmarked configuration is pinned to the reference on
machines where /root/reference does not exist."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import bindings  # noqa: E402
import synth  # noqa: E402
from make_golden import products_digest  # noqa: E402

VOX11_SCALE = 0.585


def case_frames(name):
    if name == "ai_r3":
        return [synth.figure(scale=0.626, seed=0, frame=f) for f in range(2)]
    if name == "ra_r5":
        return [synth.figure(scale=0.626, seed=0, frame=f) for f in range(4)]
    if name == "vox11":
        return [synth.figure(scale=VOX11_SCALE, seed=0, bits=11, frame=0)]
    raise KeyError(name)


def case_params(name, weight):
    """(seg params, occupancy precision) of the case"""
    if name == "ai_r3":
        return bindings.ctc_seg_params(bits=10, iterations=50, weight=weight), 4
    if name == "ra_r5":
        prm = bindings.ctc_seg_params(bits=10, iterations=50, weight=weight)
        prm.global_patch_allocation = 1
        return prm, 2
    if name == "vox11":
        return bindings.ctc_seg_params(bits=11, iterations=20, weight=weight), 4
    raise KeyError(name)


def case_bits(name):
    return 12 if name == "vox11" else 11


if __name__ == "__main__":
    ref = bindings.Reference()
    path = os.path.join(HERE, "gof_fullsize.json")
    doc = {"generator": "tests/golden/make_golden_fullsize.py",
           "source": "reference TMC2 v24.0 compiled from /root/reference (oracle/_ref), ENABLE_TBB off, single thread", "cases": {}}
    for name in (sys.argv[1:] or ["ai_r3", "ra_r5", "vox11"]):
        frames = case_frames(name)
        w = ref.weight_normal(frames[0][0], case_bits(name))
        prm, prec = case_params(name, w)
        t0 = time.perf_counter()
        gof, _ = ref.encode_gof(frames, prm, occupancy_precision=prec)
        sec = time.perf_counter() - t0
        doc["cases"][name] = {"points": [int(len(f[0])) for f in frames], "weight_normal": [float(x) for x in w],
                              "occupancy_precision": prec, "reference_seconds": round(sec, 1), "frames": products_digest(gof)}
        print(name, [len(f[0]) for f in frames], "%.1f s" % sec, [fr["patch_count"] for fr in doc["cases"][name]["frames"]], flush=True)
        with open(path, "w") as f:
            json.dump(doc, f, indent=1)
