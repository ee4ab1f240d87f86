"""(1) The 11-bit configuration of BASELINE.json configs[4] (basketball_player_vox11: geometry3dCoordinatesBitdepth 11, 2560-wide
canvas): oracle vs the reference on CPU, product vs oracle on the GPU.  (2) Size-independent properties of the product at the
full BASELINE frame size (0.83 Mpts), where the CPU oracle is too slow to be the checker: the decoder-side reconstruction
round-trips the encoder's, padding leaves occupied pixels alone, counts agree across products."""
import numpy as np
import pytest

import bindings
import synth


def vox11_gof():
    xyz, rgb = synth.figure(scale=0.2, seed=3, frame=0)
    far = (xyz.astype(np.int32) + np.array([1200, 900, 1300])).astype(np.int16)       # coordinates beyond 10 bits
    return [(far, rgb), synth.sphere(radius=25, center=1900, seed=2)]


def test_oracle_vox11_vs_reference(oracle, reference):
    frames = vox11_gof()
    assert max(int(f[0].max()) for f in frames) > 1023
    prm = bindings.ctc_seg_params(bits=11, iterations=4, weight=reference.weight_normal(frames[0][0], 12))
    ref, _ = reference.encode_gof(frames, prm)
    assert (ref[0].width, ref[0].height) == (2560, 1280)   # minimumImageWidth 2560 (cfg/sequence/*_vox11.cfg)
    assert bindings.compare_gof(oracle.encode_gof(frames, prm), ref) == []


@pytest.mark.gpu
def test_gpu_vox11_vs_oracle(oracle, product):
    frames = vox11_gof()
    w = product.weight_normal(frames[0][0], 12)
    assert np.array_equal(w, oracle.weight_normal(frames[0][0], 12))
    prm = bindings.ctc_seg_params(bits=11, iterations=4, weight=w)
    got = product.encode_gof(frames, prm)
    assert got[0].width == 2560
    assert bindings.compare_gof(got, oracle.encode_gof(frames, prm)) == []


def check_frame_properties(fr, xyz, decode=None):
    W, H, d = fr.width, fr.height, fr.data
    P = fr.patches.patches
    occ = d[1].reshape(H, W)
    om = d[2].reshape(H // 4, W // 4)
    # a16/a17: the occupancy video is the OR over 4x4 cells of the exact map; a16: one occupied pixel per valid depth0 sample
    assert np.array_equal(om, occ.reshape(H // 4, 4, W // 4, 4).max(axis=(1, 3)))
    valid = sum(int((fr.patches.depth_maps(i)[0] < 32767).sum()) for i in range(len(P)))
    assert valid == int(occ.sum()) == int(P["d0_count"].sum())
    assert all(int(p["u0"]) >= 0 and int(p["v0"]) >= 0 for p in P)
    # a18: a block maps to a patch iff its occupancy-video cells hold something
    b2p = d[3].reshape(H // 16, W // 16)
    assert np.array_equal(b2p > 0, om.reshape(H // 16, 4, W // 16, 4).max(axis=(1, 3)) > 0)
    # a19-a21: at occupied pixels D1 is at most surfaceThickness behind D0 (dilation only touches unoccupied pixels)
    g0, g1 = d[4].reshape(H, W).astype(np.int32), d[5].reshape(H, W).astype(np.int32)
    assert ((g1 - g0)[occ > 0] >= 0).all() and ((g1 - g0)[occ > 0] <= 4).all()
    # a22: one D0 point per pixel of the block-precision occupancy, pointToPixel addresses occupied cells, valid patch ids
    R = len(d[8])
    p2p = d[7].reshape(R, 3)
    up = np.repeat(np.repeat(om, 4, axis=0), 4, axis=1)
    assert (up[p2p[:, 1], p2p[:, 0]] > 0).all() and int(d[8].max()) < len(P)
    assert int((p2p[:, 2] == 0).sum()) == int(up.sum())
    if decode is not None:  # the decoder-side entry point reproduces the encoder-side reconstruction from the same images
        dec = decode(P, d[2], d[4], d[5], W, H, 4)
        assert np.array_equal(dec["xyz"].ravel(), d[6]) and np.array_equal(dec["point_to_pixel"].ravel(), d[7])
        assert np.array_equal(dec["partition"], d[8]) and np.array_equal(dec["boundary"], d[9])
    # geometry fidelity: a reconstructed D0 point on the exact occupancy is a source point (lossless maps: depth0 is a source depth)
    src = set(map(tuple, xyz.tolist()))
    rec = d[6].reshape(R, 3)
    exact = occ[p2p[:, 1], p2p[:, 0]] > 0
    sample = rec[exact & (p2p[:, 2] == 0)][::37]
    assert len(sample) and all(tuple(q) in src for q in sample.tolist())
    # a25/a26: push-pull padding changes no pixel of the block-precision occupancy; a23: colours are bytes
    for raw, pad in ((11, 13), (12, 14)):
        a, b = d[raw].reshape(3, H, W), d[pad].reshape(3, H, W)
        assert np.array_equal(a[:, up > 0], b[:, up > 0]) and int(b.max()) <= 255


def test_properties_hold_for_the_oracle(oracle):
    """the same checks the GPU test applies at full size, on a frame the oracle finishes in seconds"""
    xyz, rgb = synth.figure(scale=0.15, seed=1, frame=0)
    prm = bindings.ctc_seg_params(bits=10, iterations=5, weight=oracle.weight_normal(xyz, 11))
    check_frame_properties(oracle.encode_gof([(xyz, rgb)], prm)[0], xyz)


@pytest.mark.gpu
def test_gpu_full_size_frame_properties(product):
    """one longdress-sized frame (0.83 Mpts, I=50) through a1-a26; checks that do not need the oracle"""
    xyz, rgb = synth.figure(scale=0.626, seed=0, frame=0)
    assert len(xyz) > 800000
    prm = bindings.ctc_seg_params(bits=10, iterations=50, weight=product.weight_normal(xyz, 11))
    check_frame_properties(product.encode_gof([(xyz, rgb)], prm)[0], xyz, product.generate_point_cloud)
