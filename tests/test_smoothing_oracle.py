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
