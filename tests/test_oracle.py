"""CPU tests (no GPU): the oracle restatement against (a) the committed golden fixtures generated from the reference and
(b) the reference itself when oracle/_ref is built; plus the C-ABI surface of the product library."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import bindings
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "gof_small.json")


def _golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_matches_golden_fixture(oracle):
    m = _golden()
    with open(GOLD) as f:
        gold = json.load(f)
    frames = m.golden_frames()
    prm = m.golden_params(oracle, frames)
    assert [float(x) for x in prm.weight_normal] == gold["weight_normal"]
    got = m.products_digest(oracle.encode_gof(frames, prm))
    for f, (a, b) in enumerate(zip(got, gold["frames"])):
        for k in b:
            assert a[k] == b[k], "frame %d product %s differs from the reference fixture" % (f, k)
    ra = m.golden_ra_frames()
    got = m.products_digest(oracle.encode_gof(ra, m.golden_ra_params(oracle, ra), occupancy_precision=2))
    for f, (a, b) in enumerate(zip(got, gold["frames_random_access_r5"])):
        for k in b:
            assert a[k] == b[k], "random-access frame %d product %s differs from the reference fixture" % (f, k)
    xyz = synth.planes(n_side=20)[0]
    idx, d = oracle.knn(xyz, xyz, 16)
    assert m.digest(idx) == gold["knn16_planes20"]["idx"] and m.digest(d) == gold["knn16_planes20"]["dist"]
    assert m.digest(oracle.normals(xyz, idx, True)) == gold["normals_planes20"]


@pytest.mark.parametrize("shape", ["sphere", "planes", "random", "sheet", "specks"])
def test_oracle_knn_normals_vs_reference(shape, oracle, reference):
    xyz = {"sphere": synth.sphere(radius=16, center=40)[0], "planes": synth.planes(24)[0], "random": synth.random_cloud(1500)[0],
           "sheet": synth.double_sheet(32)[0], "specks": synth.specks()[0]}[shape]
    for k in (1, 8, 16):
        oi, od = oracle.knn(xyz, xyz[::3], k)
        ri, rd = reference.knn(xyz, xyz[::3], k)
        assert np.array_equal(oi, ri) and np.array_equal(od, rd)
    oi, _ = oracle.knn(xyz, xyz, 16)
    for orient in (False, True):
        assert np.array_equal(oracle.normals(xyz, oi, orient), reference.normals(xyz, 16, orient))
    assert np.array_equal(oracle.weight_normal(xyz, 11), reference.weight_normal(xyz, 11))
    off_o, idx_o, d_o = oracle.radius(xyz, xyz[:200], 48.0, 32767)
    off_r, idx_r, d_r = reference.radius(xyz, xyz[:200], 48.0, 32767)
    assert np.array_equal(off_o, off_r) and np.array_equal(idx_o, idx_r) and np.array_equal(d_o, d_r)


def test_oracle_fewer_points_than_k(oracle, reference):
    xyz = synth.random_cloud(9, 12, 7)[0]
    oi, od = oracle.knn(xyz, xyz, 16)
    ri, rd = reference.knn(xyz, xyz, 16)
    assert np.array_equal(oi, ri) and np.array_equal(od, rd)
    # (normals of a cloud with fewer than k points are not compared: the reference then reads stale entries of its
    #  result buffer — PCCKdTree::search resizes to k but nanoflann fills only n — so its output is not a contract)


def test_oracle_gof_vs_reference(oracle, reference):
    frames = [synth.double_sheet(n_side=48, seed=5), synth.figure(scale=0.15, seed=9, frame=2), synth.planes(n_side=30, seed=1)]
    prm = bindings.ctc_seg_params(bits=10, iterations=12, weight=reference.weight_normal(frames[0][0], 11))
    ref, _ = reference.encode_gof(frames, prm)
    assert bindings.compare_gof(oracle.encode_gof(frames, prm), ref) == []


def test_oracle_gof_precision2_vs_reference(oracle, reference):
    frames = [synth.sphere(radius=22, center=70, seed=3)]
    prm = bindings.ctc_seg_params(bits=10, iterations=5, weight=reference.weight_normal(frames[0][0], 11))
    ref, _ = reference.encode_gof(frames, prm, occupancy_precision=2)
    assert bindings.compare_gof(oracle.encode_gof(frames, prm, occupancy_precision=2), ref) == []


def test_product_library_exports_every_declared_symbol():
    """the C ABI loads and exports everything include/pccb200.h declares (no compute: there is no GPU here)"""
    if not os.path.exists(bindings.PRODUCT_SO):
        import __graft_entry__ as g
        g.build_product()
    lib = C.CDLL(bindings.PRODUCT_SO)
    with open(os.path.join(ROOT, "include", "pccb200.h")) as f:
        names = set(re.findall(r"\b(pccb200_[a-z0-9_]+)\s*\(", f.read()))
    assert len(names) >= 10
    for nme in sorted(names):
        assert hasattr(lib, nme), "libpccb200.so does not export %s" % nme


def test_product_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = C.CDLL(bindings.PRODUCT_SO)
    ctx = C.c_void_p()
    assert lib.pccb200_create(0, C.byref(ctx)) == -1  # PCCB200_ERR_NO_DEVICE: there is no CPU fallback


@pytest.mark.parametrize("name,prec,overrides", [
    ("patch_splitting", 4, dict(max_patch_size=64)),          # the a9 clamp (PCCPatchSegmenter.cpp:929-957) on clouds smaller than 1024
    ("precision1_thin", 1, dict(surface_thickness=2)),
    ("levels_and_cc", 4, dict(min_level=32, min_point_count_per_cc=8)),
    ("refine_knobs", 4, dict(lambda_refine=1.0, search_radius_refine=96)),
    ("no_orientation_no_splitting", 4, dict(normal_orientation=0, enable_patch_splitting=0)),
])
def test_oracle_gof_parameter_variations_vs_reference(name, prec, overrides, oracle, reference):
    """off-default values of the parameters pccb200_seg_params carries: the oracle follows the reference on all of them"""
    frames = [synth.figure(scale=0.2, seed=3, frame=0), synth.double_sheet(n_side=48, seed=5)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=reference.weight_normal(frames[0][0], 11))
    for k, v in overrides.items():
        setattr(prm, k, v)
    ref, _ = reference.encode_gof(frames, prm, occupancy_precision=prec)
    assert bindings.compare_gof(oracle.encode_gof(frames, prm, occupancy_precision=prec), ref) == []
