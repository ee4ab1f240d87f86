"""Generates tests/golden/gof_small.json from the REFERENCE ITSELF (oracle/_ref/libtmc2ref.so, built from /root/reference
by oracle/Makefile). Run in the build container:  python tests/golden/make_golden.py
The fixture pins sha256 digests of every hot-path product for a small deterministic GOF, so the oracle (and through it the
CUDA path) stays pinned to the reference on machines where /root/reference does not exist."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bindings  # noqa: E402
import synth  # noqa: E402


def golden_frames():
    return [synth.sphere(radius=20, center=64, seed=1), synth.double_sheet(n_side=40, seed=2), synth.specks(seed=3),
            synth.figure(scale=0.12, seed=4, frame=1)]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def products_digest(gof):
    out = []
    for fr in gof:
        d = {"width": fr.width, "height": fr.height, "patch_count": int(len(fr.patches.patches)),
             "patches": digest(fr.patches.patches), "patch_depth": digest(fr.patches.depth), "patch_occ": digest(fr.patches.occ)}
        for what, name in bindings.GOF_NAMES.items():
            d[name] = digest(fr.data[what])
            d[name + "_count"] = int(fr.data[what].size)
        out.append(d)
    return out


def golden_ra_frames():
    """random-access case (a15): a moving figure, a change of content in the middle (a second sub-context)"""
    return [synth.figure(scale=0.12, seed=4, frame=0), synth.figure(scale=0.12, seed=4, frame=1), synth.figure(scale=0.12, seed=4, frame=2),
            synth.sphere(radius=24, center=70, seed=2), synth.sphere(radius=24, center=72, seed=3)]


def golden_ra_params(ref_or_oracle, frames):
    prm = golden_params(ref_or_oracle, frames)
    prm.iteration_count_refine = 4
    prm.global_patch_allocation = 1   # cfg/condition/ctc-random-access.cfg (constrainedPack stays at its default 1)
    return prm


def golden_params(ref_or_oracle, frames):
    w = ref_or_oracle.weight_normal(frames[0][0], 11)
    return bindings.ctc_seg_params(bits=10, iterations=10, weight=w)


if __name__ == "__main__":
    ref = bindings.Reference()
    frames = golden_frames()
    prm = golden_params(ref, frames)
    gof, _ = ref.encode_gof(frames, prm)
    knn_xyz = synth.planes(n_side=20)[0]
    idx, d = ref.knn(knn_xyz, knn_xyz, 16)
    ra_frames = golden_ra_frames()
    ra_gof, _ = ref.encode_gof(ra_frames, golden_ra_params(ref, ra_frames), occupancy_precision=2)
    doc = {"generator": "tests/golden/make_golden.py", "source": "reference TMC2 v24.0 compiled from /root/reference (oracle/_ref)",
           "weight_normal": [float(x) for x in prm.weight_normal], "frames": products_digest(gof),
           "knn16_planes20": {"idx": digest(idx), "dist": digest(d)},
           "normals_planes20": digest(ref.normals(knn_xyz, 16, True)),
           "frames_random_access_r5": products_digest(ra_gof)}
    with open(os.path.join(HERE, "gof_small.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote gof_small.json:", [fr["patch_count"] for fr in doc["frames"]], "patches per frame")
