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
