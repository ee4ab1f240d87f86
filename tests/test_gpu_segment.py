"""GPU parity, rows a3..a11 of SURVEY.md §8a: orientation, initial + refined segmentation, patch list — bit-exact."""
import numpy as np
import pytest

import synth
from bindings import ctc_seg_params

pytestmark = pytest.mark.gpu

SHAPES = {
    "sphere": lambda: synth.sphere(),
    "planes": lambda: synth.planes(),
    "sheet": lambda: synth.double_sheet(),
    "specks": lambda: synth.specks(),
    "figure": lambda: synth.figure(scale=0.25),
    "figure2": lambda: synth.figure(scale=0.3, seed=5, frame=3),
}


def oracle_frame(oracle, xyz, rgb, prm):
    nbr, _ = oracle.knn(xyz, xyz, 16)
    nrm = oracle.normals(xyz, nbr, orient=True)
    p0 = oracle.initial_segmentation(nrm, np.array(list(prm.weight_normal)))
    p1 = oracle.refine_segmentation(xyz, nrm, p0, prm)
    return dict(normals=nrm, partition0=p0, partition1=p1, patches=oracle.segment_patches(xyz, rgb, nbr, p1, prm))


def compare_patches(got, want):
    assert len(got.patches) == len(want.patches), "patch count"
    for f in got.patches.dtype.names:
        assert np.array_equal(got.patches[f], want.patches[f]), "patch field %s" % f
    assert np.array_equal(got.depth, want.depth), "depth maps"
    assert np.array_equal(got.occ, want.occ), "block occupancy"


@pytest.mark.parametrize("name", list(SHAPES))
def test_oriented_normals(name, oracle, product):
    xyz = SHAPES[name]()[0]
    nbr, _ = oracle.knn(xyz, xyz, 16)
    want = oracle.normals(xyz, nbr, orient=True)
    got = product.normals(xyz, 16, orient=True)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("name", list(SHAPES))
def test_segment_frame(name, oracle, product):
    xyz, rgb = SHAPES[name]()
    prm = ctc_seg_params(bits=10, iterations=10, weight=oracle.weight_normal(xyz, 11))
    want = oracle_frame(oracle, xyz, rgb, prm)
    got = product.segment_frame(xyz, rgb, prm)
    assert np.array_equal(got["normals"].view(np.uint64), want["normals"].view(np.uint64)), "oriented normals"
    assert np.array_equal(got["partition0"], want["partition0"]), "initial segmentation"
    bad = int((got["partition1"] != want["partition1"]).sum())
    assert bad == 0, "refined segmentation differs at %d points" % bad
    compare_patches(got["patches"], want["patches"])


@pytest.mark.parametrize("name", list(SHAPES))
def test_weight_normal(name, oracle, product):
    xyz = SHAPES[name]()[0]
    assert np.array_equal(product.weight_normal(xyz, 11), oracle.weight_normal(xyz, 11))


def test_oriented_normals_full_size_frame(oracle, product):
    """one frame at the bench size (~0.83 Mpts): the sequential orientation walk must agree with the oracle everywhere"""
    xyz = synth.figure(scale=0.626, seed=0, frame=1)[0]
    nbr, _ = oracle.knn(xyz, xyz, 16)
    want = oracle.normals(xyz, nbr, orient=True)
    got = product.normals(xyz, 16, orient=True)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
