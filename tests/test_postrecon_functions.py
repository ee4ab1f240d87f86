"""The host/device functions of csrc/postrecon.cuh (the arithmetic of the not-yet-run post-reconstruction kernels), compiled by
g++ and driven sequentially (tests/postrecon_check.cpp), against the reference itself: grid-based geometry smoothing, the
YUV420 -> YUV444(16 bit) inverse conversion and convertYUV16ToRGB8."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bindings
import synth
from test_smoothing_oracle import smooth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prx():
    so = os.path.join(ROOT, "tests", "_postrecon_check.so")
    src = os.path.join(ROOT, "tests", "postrecon_check.cpp")
    hdr = os.path.join(ROOT, "mpeg-pcc-tmc2_b200", "csrc", "postrecon.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("grid,threshold", [(8, 64.0), (8, 8.0), (4, 16.0)])
def test_device_functions_geometry_smoothing_vs_reference(grid, threshold, prx, oracle, reference):
    frames = [synth.figure(scale=0.15, seed=9, frame=0), synth.double_sheet(n_side=48, seed=5)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    for fr in oracle.encode_gof(frames, prm, stop_after=3):
        xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
        want = smooth(reference.lib, "ref_smooth_geometry", xyz, bnd, part, grid, threshold)
        got = smooth(prx, "prx_smooth_geometry", xyz, bnd, part, grid, threshold)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_device_functions_colour_conversions_vs_reference(prx, reference):
    rng = np.random.default_rng(7)
    W, H = 160, 96
    y = rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8)
    outs = []
    for lib, name in ((reference.lib, "ref_yuv420_to_yuv444_16"), (prx, "prx_yuv420_to_yuv444_16")):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        o = np.zeros(3 * W * H, np.uint16)
        fn(y.ctypes.data_as(C.c_void_p), W, H, o.ctypes.data_as(C.c_void_p))
        outs.append(o)
    assert np.array_equal(outs[0], outs[1])
    yuv = rng.integers(0, 65536, (40000, 3), dtype=np.uint16)
    outs = []
    for lib, name in ((reference.lib, "ref_yuv16_to_rgb8"), (prx, "prx_yuv16_to_rgb8")):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        o = np.zeros((len(yuv), 3), np.uint8)
        fn(yuv.ctypes.data_as(C.c_void_p), len(yuv), o.ctypes.data_as(C.c_void_p))
        outs.append(o)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("spread", [257, 3])
def test_device_functions_colour_transfer_vs_reference(spread, prx, oracle, reference):
    """forwardColour / backwardColour (with the std::sort emulation) of csrc/postrecon.cuh, fed with k-NN lists and votes assembled
    here the way the kernels will assemble them, against PCCPointSet3::transferColors16bitBP run by the reference"""
    from test_smoothing_oracle import transfer
    frames = [synth.figure(scale=0.15, seed=9, frame=0)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    fr = oracle.encode_gof(frames, prm)[0]
    xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
    col16 = (fr.data[10].reshape(-1, 3).astype(np.uint16) * spread + 11).astype(np.uint16)
    sm_xyz, sm_bnd = smooth(oracle.lib._dll, "pcco_smooth_geometry", xyz, bnd, part, 8, 64.0)
    want = transfer(reference.lib, "ref_transfer_colors16_smoothed", xyz, col16, sm_xyz, col16, sm_bnd)
    moved = np.flatnonzero(sm_bnd == 3)
    # forward: 8-NN of every moved target in the source cloud
    fidx, fdist = oracle.knn(xyz, sm_xyz[moved], 8)
    refined = np.zeros((len(moved), 3), np.uint16)
    prx.prx_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    src_col = np.ascontiguousarray(col16)
    prx.prx_forward(fidx.ctypes.data_as(C.c_void_p), fdist.ctypes.data_as(C.c_void_p), 8, len(moved), src_col.ctypes.data_as(C.c_void_p), refined.ctypes.data_as(C.c_void_p))
    # backward: every sampled source point (rows in order, with repetitions) votes for its nearest target if the colours are close
    samples = fidx.ravel()
    samples = samples[samples != 0xFFFFFFFF]
    bidx, bdist = oracle.knn(sm_xyz, xyz[samples], 1)
    tgt = bidx[:, 0].astype(np.int64)
    gate = (np.abs(col16[samples].astype(np.int32) - col16[tgt].astype(np.int32)) < 40).all(axis=1)
    tgt, vdist, vcol = tgt[gate], bdist[gate, 0].astype(np.float64), np.ascontiguousarray(col16[samples][gate])
    order = np.argsort(tgt, kind="stable")                       # group by target, sampling order kept inside a group
    tgt, vdist, vcol = tgt[order], np.ascontiguousarray(vdist[order]), np.ascontiguousarray(vcol[order])
    starts, ends = np.searchsorted(tgt, moved, side="left"), np.searchsorted(tgt, moved, side="right")
    vd, vc, off = [], [], [0]                                     # CSR over the moved targets only (votes for others are never read)
    for a, b in zip(starts, ends):
        vd.append(vdist[int(a):int(b)])
        vc.append(vcol[int(a):int(b)])
        off.append(off[-1] + int(b) - int(a))
    vd = np.ascontiguousarray(np.concatenate(vd)) if vd else np.zeros(0)
    vc = np.ascontiguousarray(np.concatenate(vc)) if vc else np.zeros((0, 3), np.uint16)
    off = np.array(off, np.uint64)
    if spread == 3:
        assert max(np.diff(off.astype(np.int64))) > 16          # lists long enough to leave std::sort's insertion-sort regime
    final = np.zeros((len(moved), 3), np.uint16)
    prx.prx_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    prx.prx_backward(off.ctypes.data_as(C.c_void_p), vd.ctypes.data_as(C.c_void_p), vc.ctypes.data_as(C.c_void_p), len(moved),
                     refined.ctypes.data_as(C.c_void_p), final.ctypes.data_as(C.c_void_p))
    got = col16.copy()
    got[moved] = final
    assert np.array_equal(got, want)
