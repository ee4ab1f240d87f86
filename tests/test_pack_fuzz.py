"""Packing fuzz (a13-a15): random GOFs of patch records - no clouds, no segmentation, so thousands of patches per second - through
(1) the reference's own PCCEncoder::placeSegments (oracle/_ref), (2) the oracle's packing, (3) the product's random-access host
logic (ra_pack.hpp behind the sequential stand-in for its CUDA placement kernel). Temporal coherence, patch counts and sizes are
drawn so that the runs cover matched / unmatched patches, broken and surviving tracks, union growth, canvases above the minimum
height, rejected and accepted GPA trials and several sub-contexts per GOF."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import bindings
from test_ra_pack import shim  # noqa: F401  (fixture: builds tests/_ra_host_shim.so)


def random_gof(rng, nframes, npatch, size_hi, jitter, churn):
    """frames of (patch records, occupancy bytes); patches drift and resize from frame to frame, some die, some are born"""
    def new_patch():
        su0, sv0 = int(rng.integers(1, size_hi + 1)), int(rng.integers(1, size_hi + 1))
        return dict(view=int(rng.integers(0, 6)), u1=int(rng.integers(0, 900)), v1=int(rng.integers(0, 900)), su0=su0, sv0=sv0)
    live = [new_patch() for _ in range(npatch)]
    frames = []
    for f in range(nframes):
        recs = np.zeros(len(live), bindings.PATCH_DTYPE)
        occs, off = [], 0
        order = rng.permutation(len(live))           # creation order differs from frame to frame
        for i, k in enumerate(order):
            p = live[k]
            su = p["su0"] * 16 - int(rng.integers(0, 16))
            sv = p["sv0"] * 16 - int(rng.integers(0, 16))
            o = (rng.random(p["su0"] * p["sv0"]) < 0.8).astype(np.uint8)
            o[int(rng.integers(0, len(o)))] = 1
            r = recs[i]
            r["index"], r["view_id"], r["u1"], r["v1"], r["size_u"], r["size_v"] = i, p["view"], p["u1"], p["v1"], max(su, 1), max(sv, 1)
            r["size_u0"], r["size_v0"], r["occ_offset"], r["best_match_idx"] = p["su0"], p["sv0"], off, -1
            occs.append(o)
            off += len(o)
        frames.append((recs, np.concatenate(occs) if occs else np.zeros(0, np.uint8)))
        nxt = []
        for p in live:                                # evolve
            if rng.random() < churn:
                continue
            q = dict(p)
            q["u1"] = max(0, q["u1"] + int(rng.integers(-jitter, jitter + 1)))
            q["v1"] = max(0, q["v1"] + int(rng.integers(-jitter, jitter + 1)))
            if rng.random() < 0.3:
                q["su0"] = int(np.clip(q["su0"] + rng.integers(-1, 2), 1, size_hi))
                q["sv0"] = int(np.clip(q["sv0"] + rng.integers(-1, 2), 1, size_hi))
            nxt.append(q)
        while len(nxt) < npatch and rng.random() < 0.7:
            nxt.append(new_patch())
        live = nxt
        if not live:
            live = [new_patch()]
    return frames


def marshal(frames):
    counts = np.array([len(r) for r, _ in frames], np.int32)
    recs = np.concatenate([r for r, _ in frames])
    occ = np.concatenate([o for _, o in frames] + [np.zeros(1, np.uint8)])
    base = np.concatenate([[0], np.cumsum([len(o) for _, o in frames])]).astype(np.int64)
    return counts, recs, occ, base


def run_lib(lib, prefix, frames, ra):
    counts, recs, occ, base = marshal(frames)
    fn = getattr(lib, prefix + "pack_gof")
    fn.restype = C.c_void_p
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    for nme, rt, at in (("gof_patches", C.c_void_p, [C.c_void_p, C.c_int]), ("gof_free", None, [C.c_void_p]),
                        ("gof_dims", None, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)])):
        getattr(lib, prefix + nme).restype, getattr(lib, prefix + nme).argtypes = rt, at
    h = fn(len(frames), counts.ctypes.data_as(C.c_void_p), recs.ctypes.data_as(C.c_void_p), occ.ctypes.data_as(C.c_void_p), base.ctypes.data_as(C.c_void_p), ra, 10)
    out = []
    for f in range(len(frames)):
        w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
        getattr(lib, prefix + "gof_dims")(h, f, C.byref(w), C.byref(hh), C.byref(r))
        ps = bindings._collect_patches_borrowed(lib, prefix, getattr(lib, prefix + "gof_patches")(h, f))
        out.append((ps.patches, ps.occ, w.value, hh.value))
    getattr(lib, prefix + "gof_free")(h)
    return out


FIELDS = ("index", "view_id", "u1", "v1", "size_u", "size_v", "size_u0", "size_v0", "u0", "v0", "orientation", "best_match_idx", "is_global")


def same(a, b, what):
    for f, (x, y) in enumerate(zip(a, b)):
        assert len(x[0]) == len(y[0]), "%s: frame %d patch count" % (what, f)
        for fld in FIELDS:
            assert np.array_equal(x[0][fld], y[0][fld]), "%s: frame %d field %s" % (what, f, fld)
        assert np.array_equal(x[1], y[1]), "%s: frame %d occupancy" % (what, f)
        assert x[2:] == y[2:], "%s: frame %d canvas %s != %s" % (what, f, x[2:], y[2:])


CASES = [  # (name, GOFs, frames, patches, max blocks per side, jitter px, churn)
    ("small_stable", 6, 8, 30, 6, 6, 0.02),
    ("medium_drifting", 5, 10, 80, 8, 40, 0.08),
    ("crowded", 4, 8, 150, 10, 20, 0.05),           # close to a full 1280 x 1280 canvas
    ("crowded_mixed", 4, 10, 205, 10, 20, 0.05),    # around the limit: trials accepted for a few frames, then rejected
    ("crowded_tall", 3, 8, 220, 11, 20, 0.05),      # > 6400 blocks: canvases above 1280, every trial rejected
    ("volatile", 5, 12, 60, 9, 120, 0.35),          # tracks break constantly
    ("huge_patches", 4, 6, 25, 40, 10, 0.05),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_packing_fuzz_oracle_and_product_host_logic_vs_reference(case, oracle, reference, shim):  # noqa: F811
    name, gofs, nframes, npatch, size_hi, jitter, churn = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))   # (deterministic: str hashes are salted per process)
    stats = dict(globals=0, matched=0, tall=0, subcontexts=0)
    for g in range(gofs):
        frames = random_gof(rng, nframes, npatch, size_hi, jitter, churn)
        for ra in (0, 1):
            want = run_lib(reference.lib, "ref_", frames, ra)
            got = run_lib(oracle.lib._dll, "pcco_", frames, ra)
            same(got, want, "%s gof %d ra %d oracle vs reference" % (name, g, ra))
        # the product's host logic (random access only: all-intra packing is one kernel, covered by the GPU suite)
        counts, recs, occ, base = marshal(frames)
        assert shim.ra_shim_pack(len(frames), counts.ctypes.data_as(C.c_void_p), recs.ctypes.data_as(C.c_void_p), occ.ctypes.data_as(C.c_void_p),
                                 base.ctypes.data_as(C.c_void_p), 16, 1280, 1280) == 0
        heights = []
        for f, w in enumerate(want):
            n = shim.ra_shim_count(f)
            got = np.zeros(n, bindings.PATCH_DTYPE)
            gocc = np.zeros(shim.ra_shim_occ_bytes(f) + 1, np.uint8)
            wh = np.zeros(2, np.int64)
            shim.ra_shim_get(f, got.ctypes.data_as(C.c_void_p), gocc.ctypes.data_as(C.c_void_p), wh.ctypes.data_as(C.c_void_p))
            assert n == len(w[0]), "%s gof %d frame %d: product host logic patch count" % (name, g, f)
            for fld in FIELDS:
                assert np.array_equal(got[fld], w[0][fld]), "%s gof %d frame %d: product host logic field %s" % (name, g, f, fld)
            assert np.array_equal(gocc[:-1], w[1])
            heights.append(int(wh[1]))
            stats["globals"] += int(w[0]["is_global"].sum())
            stats["matched"] += int((w[0]["best_match_idx"] >= 0).sum())
            stats["subcontexts"] += int(f > 0 and len(w[0]) > 0 and (w[0]["best_match_idx"] >= 0).sum() == 0)
        assert max(1280, -(-max(heights) // 64) * 64) == want[0][3]
        stats["tall"] += int(want[0][3] > 1280)
    assert stats["globals"] > 0
    if name == "crowded_tall":     # every frame its own sub-context (the first frame of a sub-context keeps no match index)
        assert stats["tall"] > 0 and stats["subcontexts"] > 0
    if name in ("small_stable", "medium_drifting", "crowded", "crowded_mixed", "volatile"):
        assert stats["matched"] > 0
    if name == "crowded_mixed":    # sub-contexts of several frames AND sub-context breaks in the same GOF
        assert stats["subcontexts"] > 0
