"""GPU parity of the post-reconstruction chain (csrc/postrecon.cu, entry points declared in include/pccb200.h) against
the oracle, which is pinned against the reference on the CPU (tests/test_smoothing_oracle.py); the decoder-side binding
(integration/pccb200_shim.cpp decodeFrame) against the reference's own generatePointCloud; and the random-cloud fuzz of
tests/test_oracle_fuzz.py with the product in place of the oracle."""
import ctypes as C

import numpy as np
import pytest

import bindings
import synth
from test_smoothing_oracle import smooth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid,threshold", [(8, 64.0), (8, 8.0), (4, 16.0)])
def test_gpu_geometry_smoothing_vs_oracle(grid, threshold, oracle, product):
    frames = [synth.figure(scale=0.15, seed=9, frame=0), synth.double_sheet(n_side=48, seed=5)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    fn = product.lib.pccb200_smooth_geometry
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double]
    for fr in oracle.encode_gof(frames, prm, stop_after=3):
        xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
        want = smooth(oracle.lib._dll, "pcco_smooth_geometry", xyz, bnd, part, grid, threshold)
        x, b = np.ascontiguousarray(xyz, np.int16).copy(), np.ascontiguousarray(bnd, np.uint16).copy()
        p = np.ascontiguousarray(part, np.uint32)
        assert fn(product.ctx, x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), len(b), grid, threshold) == 0
        assert np.array_equal(x, want[0]) and np.array_equal(b, want[1])


def test_gpu_colour_conversions_vs_oracle(oracle, product):
    rng = np.random.default_rng(7)
    W, H = 1280, 1344
    y = rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8)
    want = np.zeros(3 * W * H, np.uint16)
    f = oracle.lib._dll.pcco_yuv420_to_yuv444_16
    f.restype, f.argtypes = None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    f(y.ctypes.data_as(C.c_void_p), W, H, want.ctypes.data_as(C.c_void_p))
    got = np.zeros_like(want)
    g = product.lib.pccb200_yuv420_to_yuv444_16
    g.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    assert g(product.ctx, y.ctypes.data_as(C.c_void_p), W, H, got.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(got, want)
    yuv = rng.integers(0, 65536, (300000, 3), dtype=np.uint16)
    want = np.zeros((len(yuv), 3), np.uint8)
    f = oracle.lib._dll.pcco_yuv16_to_rgb8
    f.restype, f.argtypes = None, [C.c_void_p, C.c_size_t, C.c_void_p]
    f(yuv.ctypes.data_as(C.c_void_p), len(yuv), want.ctypes.data_as(C.c_void_p))
    got = np.zeros_like(want)
    g = product.lib.pccb200_yuv16_to_rgb8
    g.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    assert g(product.ctx, yuv.ctypes.data_as(C.c_void_p), len(yuv), got.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("spread", [257, 3])
def test_gpu_colour_transfer_onto_smoothed_cloud_vs_oracle(spread, oracle, product):
    from test_smoothing_oracle import transfer
    frames = [synth.figure(scale=0.15, seed=9, frame=0)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=oracle.weight_normal(frames[0][0], 11))
    fr = oracle.encode_gof(frames, prm)[0]
    xyz, bnd, part = fr.data[6].reshape(-1, 3), fr.data[9], fr.data[8]
    col16 = (fr.data[10].reshape(-1, 3).astype(np.uint16) * spread + 11).astype(np.uint16)
    sm_xyz, sm_bnd = smooth(oracle.lib._dll, "pcco_smooth_geometry", xyz, bnd, part, 8, 64.0)
    want = transfer(oracle.lib._dll, "pcco_transfer_colors16_smoothed", xyz, col16, sm_xyz, col16, sm_bnd)
    fn = product.lib.pccb200_transfer_colors16_smoothed
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    sx, sc = np.ascontiguousarray(xyz, np.int16), np.ascontiguousarray(col16)
    tx, tc, tb = np.ascontiguousarray(sm_xyz, np.int16), col16.copy(), np.ascontiguousarray(sm_bnd, np.uint16)
    assert fn(product.ctx, sx.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), len(sx), tx.ctypes.data_as(C.c_void_p),
              tc.ctypes.data_as(C.c_void_p), tb.ctypes.data_as(C.c_void_p), len(tx)) == 0
    assert np.array_equal(tc, want)


def test_gpu_decoder_side_binding_vs_reference():
    """integration/pccb200_shim.cpp decodeFrame (what PCCDecoder::decode would call instead of generatePointCloud): the reference's
    own encoder stages produce patches + occupancy + geometry frames, the reconstruction then comes from the decoder-side binding"""
    shim = bindings.Shim()
    frames = [synth.figure(scale=0.12, seed=4, frame=f) for f in range(2)] + [synth.double_sheet(n_side=32, seed=2)]
    prm = bindings.ctc_seg_params(bits=10, iterations=4, weight=(1.0, 1.0, 1.0))
    want, _ = shim.ref.encode_gof(frames, prm, stop_after=3)
    got, code = shim.decode_gof(frames, prm)
    assert code == 0
    for a, b in zip(got, want):
        for what in (6, 7, 8, 9):    # positions, pointToPixel, partition, boundary point types
            assert np.array_equal(a.data[what], b.data[what]), bindings.GOF_NAMES[what]


@pytest.mark.parametrize("kind", ["blob", "shell", "planes", "dust"])
def test_gpu_product_vs_oracle_on_random_clouds(kind, oracle, product):
    """the unstructured shapes of tests/test_oracle_fuzz.py with the product in place of the oracle (both packing modes)"""
    from test_oracle_fuzz import cloud
    rng = np.random.default_rng({"blob": 11, "shell": 12, "planes": 13, "dust": 14}[kind])
    frames = [cloud(kind, rng), cloud(kind, rng)]
    for ra, prec in ((0, 4), (1, 2)):
        prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=product.weight_normal(frames[0][0], 11))
        prm.global_patch_allocation = ra
        assert bindings.compare_gof(product.encode_gof(frames, prm, occupancy_precision=prec), oracle.encode_gof(frames, prm, occupancy_precision=prec)) == []
