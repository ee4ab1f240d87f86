"""Random-access packing (spatialConsistencyPackFlexible + data-adaptive global patch allocation).

CPU: (1) the oracle restatement against the reference itself (oracle/_ref) on GOFs that exercise matched / unmatched patches,
several sub-contexts, broken tracks and canvases above the minimum height; (2) the HOST logic of the product's packer
(mpeg-pcc-tmc2_b200/csrc/ra_pack.hpp) against the oracle, driven through a sequential stand-in for the CUDA placement kernel
(tests/ra_host_shim.cpp, test-only).  GPU (-m gpu): the product's GOF entry point with global_patch_allocation = 1 against
the oracle, all products."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bindings
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fig(f, s=0.15, seed=9):
    return synth.figure(scale=s, seed=seed, frame=f)


def gof(name):
    if name == "mixed":      # tracks break at every change of shape: several sub-contexts, with and without a reference frame
        return [fig(0), fig(1), synth.sphere(radius=30, center=80, seed=1), synth.sphere(radius=30, center=82, seed=2), fig(2), fig(3),
                synth.planes(n_side=30, seed=1), fig(4)]
    if name == "jumpy":      # large motion: matches survive, unions grow
        return [fig(0), fig(7), fig(14), fig(21), fig(3, 0.15, 4), fig(4, 0.15, 4)]
    if name == "stack":      # >100 block-sized sheets: the packing exceeds the minimum image height
        return [synth.sheet_stack(11, f, spacing=6, gap=28) for f in range(4)]
    if name == "single":
        return [fig(0)]
    if name == "empty_mid":  # an empty frame is legal (PCCEncoder.cpp:3688, 4772-4774): it ends the sub-context and starts its own
        empty = (np.zeros((0, 3), np.int16), np.zeros((0, 3), np.uint8))
        return [fig(0, 0.12), fig(1, 0.12), empty, fig(2, 0.12), fig(3, 0.12)]
    if name == "tiny":       # a 40-point frame (one patch) between two ordinary ones
        return [fig(0, 0.12), synth.random_cloud(40, 8, 1), fig(1, 0.12)]
    raise KeyError(name)


def ra_params(weight, iterations=3):
    prm = bindings.ctc_seg_params(bits=10, iterations=iterations, weight=weight)
    prm.global_patch_allocation = 1
    return prm


@pytest.mark.parametrize("name", ["mixed", "jumpy", "stack", "single", "empty_mid", "tiny"])
def test_oracle_ra_packing_vs_reference(name, oracle, reference):
    frames = gof(name)
    prm = ra_params(reference.weight_normal(frames[0][0], 11))
    ref, _ = reference.encode_gof(frames, prm, occupancy_precision=2, stop_after=1)
    got = oracle.encode_gof(frames, prm, occupancy_precision=2, stop_after=1)
    assert bindings.compare_gof(got, ref) == []
    if name == "mixed":   # the case must really cover global patches, matches and more than one sub-context
        assert sum(int(f.patches.patches["is_global"].sum()) for f in ref) > 50
        assert sum(1 for f in ref[1:] if (f.patches.patches["best_match_idx"] >= 0).sum() == 0) >= 2


def test_oracle_ra_full_products_vs_reference(oracle, reference):
    frames = [fig(0), fig(1), fig(2)]
    prm = ra_params(reference.weight_normal(frames[0][0], 11))
    ref, _ = reference.encode_gof(frames, prm, occupancy_precision=2)
    assert bindings.compare_gof(oracle.encode_gof(frames, prm, occupancy_precision=2), ref) == []


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(ROOT, "tests", "_ra_host_shim.so")
    src = os.path.join(ROOT, "tests", "ra_host_shim.cpp")
    hdr = os.path.join(ROOT, "mpeg-pcc-tmc2_b200", "csrc", "ra_pack.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.ra_shim_occ_bytes.restype = C.c_size_t
    return lib


@pytest.mark.parametrize("name", ["mixed", "jumpy", "stack", "single", "empty_mid", "tiny"])
def test_product_ra_host_logic_vs_oracle(name, oracle, shim):
    frames = gof(name)
    prm = ra_params(oracle.weight_normal(frames[0][0], 11), iterations=2)
    want = oracle.encode_gof(frames, prm, occupancy_precision=2, stop_after=1)
    sets = [oracle.segment_frame_patches(x, c, prm) for x, c in frames]
    counts = np.array([len(s.patches) for s in sets], np.int32)
    patches = np.concatenate([s.patches for s in sets])
    occ = np.concatenate([s.occ for s in sets] + [np.zeros(1, np.uint8)])
    base = np.concatenate([[0], np.cumsum([len(s.occ) for s in sets])]).astype(np.int64)
    rc = shim.ra_shim_pack(len(frames), counts.ctypes.data_as(C.c_void_p), patches.ctypes.data_as(C.c_void_p), occ.ctypes.data_as(C.c_void_p),
                           base.ctypes.data_as(C.c_void_p), 16, 1280, 1280)
    assert rc == 0
    heights = []
    for f, w in enumerate(want):
        n = shim.ra_shim_count(f)
        got = np.zeros(n, bindings.PATCH_DTYPE)
        gocc = np.zeros(shim.ra_shim_occ_bytes(f) + 1, np.uint8)
        wh = np.zeros(2, np.int64)
        shim.ra_shim_get(f, got.ctypes.data_as(C.c_void_p), gocc.ctypes.data_as(C.c_void_p), wh.ctypes.data_as(C.c_void_p))
        wp = w.patches.patches
        assert n == len(wp), "frame %d patch count" % f
        for fld in ("index", "view_id", "u1", "v1", "size_u", "size_v", "size_u0", "size_v0", "u0", "v0", "orientation", "best_match_idx", "is_global"):
            assert np.array_equal(got[fld], wp[fld]), "frame %d field %s" % (f, fld)
        assert np.array_equal(gocc[:-1], w.patches.occ), "frame %d occupancy" % f
        heights.append(int(wh[1]))
    # the GOF canvas is the maximum over the frames, at least the minimum image size, rounded up to 64
    assert max(1280, -(-max(heights) // 64) * 64) == want[0].height
    if name == "stack":
        assert want[0].height > 1280


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mixed", "jumpy", "stack", "single", "empty_mid", "tiny"])
def test_gpu_ra_gof_vs_oracle(name, oracle, product):
    frames = gof(name)
    w = product.weight_normal(frames[0][0], 11)
    assert np.array_equal(w, oracle.weight_normal(frames[0][0], 11))
    prm = ra_params(w, iterations=2)
    stop = 1 if name == "stack" else 0
    got = product.encode_gof(frames, prm, occupancy_precision=2, stop_after=stop)
    want = oracle.encode_gof(frames, prm, occupancy_precision=2, stop_after=stop)
    assert bindings.compare_gof(got, want) == []


@pytest.mark.gpu
def test_gpu_ra_gof_vs_reference_fixture(product):
    """the committed fixture was generated from the reference itself (tests/golden/make_golden.py): every product digest must match"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with open(os.path.join(ROOT, "tests", "golden", "gof_small.json")) as f:
        gold = json.load(f)["frames_random_access_r5"]
    frames = m.golden_ra_frames()
    got = m.products_digest(product.encode_gof(frames, m.golden_ra_params(product, frames), occupancy_precision=2))
    for f, (a, b) in enumerate(zip(got, gold)):
        for k in b:
            assert a[k] == b[k], "random-access frame %d product %s differs from the reference fixture" % (f, k)


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [2, 3])
def test_gpu_sharded_random_access_gof_equals_unsharded(shards, product):
    """SURVEY 8e, random access: the frames of a GOF held by several ranks (here: several library contexts on one GPU), segmentation
    per shard (stop_after = 5), ONE exchange of patch records + block occupancies, the deterministic packing replicated on every
    shard (pccb200_gof_pack_ra), image formation per shard - every product must equal the unsharded GOF's"""
    frames = [synth.figure(scale=0.14, seed=4, frame=f) for f in range(5)] + [synth.sphere(radius=22, center=70, seed=2)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=product.weight_normal(frames[0][0], 11))
    prm.global_patch_allocation = 1
    whole = product.encode_gof(frames, prm, occupancy_precision=2)
    ctxs = [bindings.Product(0) for _ in range(shards)]
    owner = [f % shards for f in range(len(frames))]                       # frame f -> shard f mod shards
    local = [[f for f in range(len(frames)) if owner[f] == r] for r in range(shards)]
    gofs = [bindings.ProductGof(ctxs[r], [frames[f] for f in local[r]], prm, 2, stop_after=5) for r in range(shards)]
    records = [None] * len(frames)                                         # the all-gather: every shard learns every frame's records
    for r in range(shards):
        for i, f in enumerate(local[r]):
            records[f] = gofs[r].patch_records(i)
    for r in range(shards):
        gofs[r].pack_ra(records, [local[r].index(f) if owner[f] == r else -1 for f in range(len(frames))])
    dims = {gofs[r].dims(0)[:2] for r in range(shards)}
    assert dims == {(whole[0].width, whole[0].height)}, "every shard must arrive at the GOF-wide canvas"
    for r in range(shards):
        W, H, _ = gofs[r].dims(0)
        gofs[r].resume(W, H, 0)
        for i, f in enumerate(local[r]):
            got = gofs[r].patch_list(i)
            for fld in got.patches.dtype.names:
                assert np.array_equal(got.patches[fld], whole[f].patches.patches[fld]), "frame %d patch field %s" % (f, fld)
            assert np.array_equal(got.depth, whole[f].patches.depth) and np.array_equal(got.occ, whole[f].patches.occ)
            for what, name in bindings.GOF_NAMES.items():
                assert np.array_equal(gofs[r].fetch(i, what), whole[f].data[what]), "frame %d %s (shard %d)" % (f, name, r)
        gofs[r].free()
        ctxs[r].close()
