"""§8f-1, first stage of the post-reconstruction chain: the oracle's grid-based geometry smoothing
(PCCCodec::smoothPointCloudPostprocess, PccLibCommon/source/PCCCodec.cpp:54-150, 982-1106) against the reference itself on
reconstructed clouds (positions, boundary point types, patch index per point as generatePointCloud leaves them). Oracle only:
the CUDA stage comes next; this pins the checker first."""
import ctypes as C

import numpy as np
import pytest

import bindings
import synth


def smooth(lib, name, xyz, boundary, partition, grid, threshold):
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double]
    x, b = np.ascontiguousarray(xyz, np.int16).copy(), np.ascontiguousarray(boundary, np.uint16).copy()
    p = np.ascontiguousarray(partition, np.uint32)
    fn(x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), len(b), grid, threshold)
    return x, b


@pytest.mark.parametrize("grid,threshold", [(8, 64.0), (8, 8.0), (4, 16.0), (16, 64.0)])
def test_oracle_geometry_smoothing_vs_reference(grid, threshold, oracle, reference):
    frames = [synth.figure(scale=0.15, seed=9, frame=0), synth.double_sheet(n_side=48, seed=5), synth.sphere(radius=5, center=6, seed=1)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    moved = 0
    for fr in oracle.encode_gof(frames, prm, stop_after=3):
        xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
        want = smooth(reference.lib, "ref_smooth_geometry", xyz, bnd, part, grid, threshold)
        got = smooth(oracle.lib._dll, "pcco_smooth_geometry", xyz, bnd, part, grid, threshold)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        moved += int((want[1] == 3).sum())
    if threshold <= 16.0:
        assert moved > 0   # the case must actually smooth something


def transfer(lib, name, src_xyz, src_col, tgt_xyz, tgt_col, tgt_bnd):
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    sx, sc = np.ascontiguousarray(src_xyz, np.int16), np.ascontiguousarray(src_col, np.uint16)
    tx, tc, tb = np.ascontiguousarray(tgt_xyz, np.int16), np.ascontiguousarray(tgt_col, np.uint16).copy(), np.ascontiguousarray(tgt_bnd, np.uint16)
    fn(sx.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), len(sx), tx.ctypes.data_as(C.c_void_p), tc.ctypes.data_as(C.c_void_p),
       tb.ctypes.data_as(C.c_void_p), len(tx))
    return tc


@pytest.mark.parametrize("spread", [257, 3])   # 16-bit colours far apart (the < 40 gate rejects most votes) / close together (most pass)
def test_oracle_colour_transfer_onto_smoothed_cloud_vs_reference(spread, oracle, reference):
    """second stage of the chain: PCCPointSet3::transferColors16bitBP as encode / decode call it (PCCPointSet.cpp:1126-1485)"""
    frames = [synth.figure(scale=0.15, seed=9, frame=0)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    fr = oracle.encode_gof(frames, prm)[0]
    xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
    col16 = fr.data[10].reshape(-1, 3).astype(np.uint16) * spread + 11
    sm_xyz, sm_bnd = smooth(oracle.lib._dll, "pcco_smooth_geometry", xyz, bnd, part, 8, 64.0)
    assert int((sm_bnd == 3).sum()) > 1000
    want = transfer(reference.lib, "ref_transfer_colors16_smoothed", xyz, col16, sm_xyz, col16, sm_bnd)
    got = transfer(oracle.lib._dll, "pcco_transfer_colors16_smoothed", xyz, col16, sm_xyz, col16, sm_bnd)
    assert np.array_equal(got, want)
    assert np.array_equal(want[sm_bnd != 3], col16[sm_bnd != 3]) and int((want != col16).any(axis=1).sum()) > 500


def test_oracle_inverse_colour_conversion_and_final_rgb_vs_reference(oracle, reference):
    """the ends of the chain: decoded YUV 4:2:0 -> 16-bit YUV 4:4:4 ("YUV420ToYUV444_8_0") and convertYUV16ToRGB8"""
    rng = np.random.default_rng(5)
    W, H = 128, 96
    smooth_img = (np.add.outer(np.arange(H), np.arange(W)) % 256).astype(np.uint8)
    for y in (rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8), np.concatenate([smooth_img.ravel(), smooth_img[::2, ::2].ravel(), 255 - smooth_img[::2, ::2].ravel()])):
        outs = []
        for lib, name in ((reference.lib, "ref_yuv420_to_yuv444_16"), (oracle.lib._dll, "pcco_yuv420_to_yuv444_16")):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
            o = np.zeros(3 * W * H, np.uint16)
            src = np.ascontiguousarray(y)
            fn(src.ctypes.data_as(C.c_void_p), W, H, o.ctypes.data_as(C.c_void_p))
            outs.append(o)
        assert np.array_equal(outs[0], outs[1])
    yuv = rng.integers(0, 65536, (50000, 3), dtype=np.uint16)
    yuv[:100] = [[0, 0, 0]] * 50 + [[65535, 65535, 65535]] * 50
    outs = []
    for lib, name in ((reference.lib, "ref_yuv16_to_rgb8"), (oracle.lib._dll, "pcco_yuv16_to_rgb8")):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        o = np.zeros((len(yuv), 3), np.uint8)
        fn(yuv.ctypes.data_as(C.c_void_p), len(yuv), o.ctypes.data_as(C.c_void_p))
        outs.append(o)
    assert np.array_equal(outs[0], outs[1])
