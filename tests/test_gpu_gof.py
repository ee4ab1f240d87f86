"""GPU parity, rows a13..a26 of SURVEY.md §8a through the GOF entry point: packing, occupancy / geometry images,
block dilation, generatePointCloud, colour transfer, attribute images and push-pull padding — all bit-exact vs the oracle."""
import numpy as np
import pytest

import bindings
import synth
from bindings import compare_gof, ctc_seg_params

pytestmark = pytest.mark.gpu

GOFS = {
    "figures": lambda: [synth.figure(scale=0.25, frame=f) for f in range(2)],
    "mixed": lambda: [synth.sphere(), synth.double_sheet(), synth.specks(), synth.planes()],
    "with_empty_frame": lambda: [synth.sphere(radius=20, center=60), (np.zeros((0, 3), np.int16), np.zeros((0, 3), np.uint8))],
}


@pytest.mark.parametrize("name", list(GOFS))
def test_encode_gof_all_products(name, oracle, product):
    frames = GOFS[name]()
    prm = ctc_seg_params(bits=10, iterations=10, weight=product.weight_normal(frames[0][0], 11))
    want = oracle.encode_gof(frames, prm)
    got = product.encode_gof(frames, prm)
    assert compare_gof(got, want) == []


def test_encode_gof_precision2(oracle, product):
    frames = [synth.sphere(radius=22, center=70, seed=3), synth.double_sheet(n_side=40, seed=2)]
    prm = ctc_seg_params(bits=10, iterations=5, weight=product.weight_normal(frames[0][0], 11))
    assert compare_gof(product.encode_gof(frames, prm, occupancy_precision=2), oracle.encode_gof(frames, prm, occupancy_precision=2)) == []


def test_encode_gof_vs_reference_if_built(product):
    if not bindings.Reference.available():
        pytest.skip("oracle/_ref not present")
    ref = bindings.Reference()
    frames = [synth.figure(scale=0.2, seed=2, frame=1), synth.specks(seed=4)]
    prm = ctc_seg_params(bits=10, iterations=10, weight=product.weight_normal(frames[0][0], 11))
    want, _ = ref.encode_gof(frames, prm)
    assert compare_gof(product.encode_gof(frames, prm), want) == []


def test_decoder_side_generate_point_cloud(oracle, product):
    """PCCCodec::generatePointCloud as PCCDecoder calls it: patch records + decoded planes in, cloud out."""
    frames = [synth.figure(scale=0.2, seed=7, frame=0), synth.double_sheet(n_side=48, seed=1)]
    prm = ctc_seg_params(bits=10, iterations=8, weight=product.weight_normal(frames[0][0], 11))
    want = oracle.encode_gof(frames, prm, stop_after=3)
    for fr in want:
        got = product.generate_point_cloud(fr.patches.patches, fr[bindings.GOF_OM_VIDEO], fr[bindings.GOF_GEO0], fr[bindings.GOF_GEO1], fr.width, fr.height)
        assert np.array_equal(got["xyz"].ravel(), fr[bindings.GOF_REC_XYZ])
        assert np.array_equal(got["point_to_pixel"].ravel(), fr[bindings.GOF_POINT_TO_PIXEL])
        assert np.array_equal(got["partition"], fr[bindings.GOF_REC_PARTITION])
        assert np.array_equal(got["boundary"], fr[bindings.GOF_REC_BOUNDARY])


def test_lossy_codec_round_trip_protocol(oracle, product):
    """stop after the geometry images, hand 'decoded' planes back, resume: identity planes reproduce the one-shot result and
    perturbed planes reproduce what the decoder-side entry point reconstructs from the same planes."""
    frames = [synth.figure(scale=0.2, seed=3, frame=1)]
    prm = ctc_seg_params(bits=10, iterations=8, weight=product.weight_normal(frames[0][0], 11))
    one_shot = product.encode_gof(frames, prm)
    g = bindings.ProductGof(product, frames, prm, 4)
    W, H, _ = g.dims(0)
    g.resume(W, H, 2)
    om, g0, g1 = g.fetch(0, bindings.GOF_OM_VIDEO), g.fetch(0, bindings.GOF_GEO0), g.fetch(0, bindings.GOF_GEO1)
    assert np.array_equal(g0, one_shot[0][bindings.GOF_GEO0]) and np.array_equal(om, one_shot[0][bindings.GOF_OM_VIDEO])
    noisy0 = np.clip(g0.astype(np.int32) + (np.arange(g0.size) % 3 == 0), 0, 255).astype(np.uint16)   # a 'lossy' D0
    noisy1 = np.maximum(noisy0, g1)
    g.set_decoded(0, om, noisy0, noisy1)
    g.resume(W, H, 0)
    rec = g.fetch(0, bindings.GOF_REC_XYZ)
    p2p = g.fetch(0, bindings.GOF_POINT_TO_PIXEL)
    g.free()
    ref = product.generate_point_cloud(one_shot[0].patches.patches, om, noisy0, noisy1, W, H)
    assert np.array_equal(rec, ref["xyz"].ravel()) and np.array_equal(p2p, ref["point_to_pixel"].ravel())
    assert not np.array_equal(rec, one_shot[0][bindings.GOF_REC_XYZ])


def test_encode_gof_vs_reference_fixture(product):
    """tests/golden/gof_small.json holds digests of every product as computed by the reference itself (make_golden.py)"""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(root, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with open(os.path.join(root, "tests", "golden", "gof_small.json")) as f:
        gold = json.load(f)
    frames = m.golden_frames()
    prm = m.golden_params(product, frames)
    assert [float(x) for x in prm.weight_normal] == gold["weight_normal"]
    got = m.products_digest(product.encode_gof(frames, prm))
    for f, (a, b) in enumerate(zip(got, gold["frames"])):
        for k in b:
            assert a[k] == b[k], "frame %d product %s differs from the reference fixture" % (f, k)


@pytest.mark.parametrize("name,prec,overrides", [
    ("patch_splitting", 4, dict(max_patch_size=64)),   # the a9 clamp, which clouds smaller than 1024 never reach at the default
    ("precision1_thin_levels_cc", 1, dict(surface_thickness=2, min_level=32, min_point_count_per_cc=8)),
    ("refine_knobs_no_orientation_no_splitting", 4, dict(lambda_refine=1.0, search_radius_refine=96, normal_orientation=0, enable_patch_splitting=0)),
])
def test_encode_gof_parameter_variations(name, prec, overrides, oracle, product):
    """off-default values of the parameters pccb200_seg_params carries (the oracle follows the reference on them: tests/test_oracle.py)"""
    frames = [synth.figure(scale=0.12, seed=3, frame=0), synth.double_sheet(n_side=32, seed=5)]
    prm = ctc_seg_params(bits=10, iterations=3, weight=product.weight_normal(frames[0][0], 11))
    for k, v in overrides.items():
        setattr(prm, k, v)
    assert compare_gof(product.encode_gof(frames, prm, occupancy_precision=prec), oracle.encode_gof(frames, prm, occupancy_precision=prec)) == []


def test_reform_on_larger_canvas(oracle, product):
    """bench.py's sharded protocol: a rank forms the images on its local canvas size and, if the GOF-wide maximum turns out larger,
    again on that one (pccb200_gof_resume on a finished GOF) - the second result must equal a one-shot run on the larger canvas"""
    frames = [synth.figure(scale=0.2, seed=5, frame=0), synth.double_sheet(n_side=40, seed=3)]
    prm = ctc_seg_params(bits=10, iterations=6, weight=product.weight_normal(frames[0][0], 11))
    g = bindings.ProductGof(product, frames, prm, 4)
    W, H, _ = g.dims(0)
    g.resume(W, H, 0)
    small = [g.fetch(f, bindings.GOF_GEO0) for f in range(2)]
    g.resume(W, H + 128, 0)
    assert g.dims(0)[:2] == (W, H + 128)
    want = oracle.encode_gof(frames, prm, canvas=(W, H + 128))
    for f in range(2):
        for what in bindings.GOF_NAMES:
            assert np.array_equal(g.fetch(f, what), want[f].data[what]), "frame %d %s after re-forming" % (f, bindings.GOF_NAMES[what])
        assert small[f].size < want[f].data[bindings.GOF_GEO0].size
    g.free()


def test_unsupported_parameters_are_rejected(product):
    """anything outside the implemented path is PCCB200_ERR_UNSUPPORTED, never a silently different result"""
    frames = [synth.sphere(radius=12, center=40)]
    for field, value in (("voxel_dim_refine", 2), ("voxel_dim_refine", 1), ("voxel_dim_refine", 8), ("search_radius_refine", 4096),
                         ("nn_normal_estimation", 8), ("map_count_minus1", 0)):
        prm = ctc_seg_params(bits=10, iterations=2, weight=(1.0, 1.0, 1.0))
        setattr(prm, field, value)
        with pytest.raises(RuntimeError, match="error -6"):
            product.encode_gof(frames, prm)


def test_decoder_rejects_patches_outside_the_canvas(oracle, product):
    """the decoder-side entry point is fed from a bitstream: negative or oversized placements must be refused, not rasterised"""
    frames = [synth.sphere(radius=16, center=50, seed=1)]
    prm = ctc_seg_params(bits=10, iterations=3, weight=product.weight_normal(frames[0][0], 11))
    fr = oracle.encode_gof(frames, prm, stop_after=3)[0]
    args = (fr[bindings.GOF_OM_VIDEO], fr[bindings.GOF_GEO0], fr[bindings.GOF_GEO1], fr.width, fr.height)
    good = product.generate_point_cloud(fr.patches.patches, *args)
    assert np.array_equal(good["xyz"].ravel(), fr[bindings.GOF_REC_XYZ])
    for field, value in (("u0", -1), ("v0", -3), ("size_u0", 0), ("size_v0", -2), ("u0", fr.width // 16), ("v0", fr.height // 16)):
        bad = fr.patches.patches.copy()
        bad[field][0] = value
        with pytest.raises(RuntimeError):
            product.generate_point_cloud(bad, *args)
    with pytest.raises(RuntimeError):
        product.generate_point_cloud(fr.patches.patches, fr[bindings.GOF_OM_VIDEO], fr[bindings.GOF_GEO0], fr[bindings.GOF_GEO1], 0, fr.height)


def test_geometry_luma_as_bytes(product):
    """PCCB200_GOF_GEO0_LUMA8 / GEO1_LUMA8: the geometry planes narrowed on the device for an 8-bit codec (half the D2H bytes)"""
    frames = [synth.figure(scale=0.2, seed=6, frame=0)]
    prm = ctc_seg_params(bits=10, iterations=4, weight=product.weight_normal(frames[0][0], 11))
    g = bindings.ProductGof(product, frames, prm, 4)
    assert g.count(0, bindings.GOF_GEO0_LUMA8) == 0          # not before the geometry images exist
    W, H, _ = g.dims(0)
    g.resume(W, H, 2)
    for wide, narrow in ((bindings.GOF_GEO0, bindings.GOF_GEO0_LUMA8), (bindings.GOF_GEO1, bindings.GOF_GEO1_LUMA8)):
        a, b = g.fetch(0, wide), g.fetch(0, narrow)
        assert b.dtype == np.uint8 and b.size == W * H and int(a.max()) <= 255 and np.array_equal(a.astype(np.uint8), b)
    g.free()
