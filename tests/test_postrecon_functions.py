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
