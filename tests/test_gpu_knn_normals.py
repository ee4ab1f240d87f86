"""GPU parity, rows a1/a2/a5 of SURVEY.md §8a: kd-tree order, k-NN lists, PCA normals — bit-exact vs the oracle."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

SHAPES = {
    "sphere": lambda: synth.sphere(),
    "planes": lambda: synth.planes(),
    "random": lambda: synth.random_cloud(),
    "sheet": lambda: synth.double_sheet(),
    "specks": lambda: synth.specks(),
    "figure": lambda: synth.figure(scale=0.25),
    "tiny7": lambda: (synth.random_cloud(7, 16, 3)[0], None),
    "tiny11": lambda: (synth.random_cloud(11, 16, 4)[0], None),
}


@pytest.mark.parametrize("name", list(SHAPES))
def test_tree_order_and_knn16(name, oracle, product):
    xyz = SHAPES[name]()[0]
    assert np.array_equal(product.vind(xyz), oracle.vind(xyz)), "leaf order (vind) differs from nanoflann's"
    gi, gd = product.knn(xyz, None, 16)
    oi, od = oracle.knn(xyz, xyz, 16)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)


@pytest.mark.parametrize("k", [1, 8])
def test_knn_foreign_queries(k, oracle, product):
    xyz = synth.figure(scale=0.2)[0]
    q = synth.figure(scale=0.2, seed=3, frame=2)[0][:5000]
    gi, gd = product.knn(xyz, q, k)
    oi, od = oracle.knn(xyz, q, k)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


@pytest.mark.parametrize("name", ["sphere", "planes", "sheet", "figure"])
def test_normals_unoriented(name, oracle, product):
    xyz = SHAPES[name]()[0]
    nbr, _ = oracle.knn(xyz, xyz, 16)
    want = oracle.normals(xyz, nbr, orient=False)
    got = product.normals(xyz, 16, orient=False)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), "fp64 normals must match bit for bit"
