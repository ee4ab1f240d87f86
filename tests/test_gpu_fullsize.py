"""GPU parity AT THE BENCHMARKED SIZES against the reference itself: tests/golden/gof_fullsize.json holds sha256 digests of every
hot-path product (patch lists incl. depth / occupancy arenas, a13-a26 images, reconstruction, colours) computed by the compiled
reference (tests/golden/make_golden_fullsize.py) for
  ai_r3  two 0.83 Mpts frames of the default bench workload, CTC all-intra r3, I=50     (BASELINE.json configs[1])
  ra_r5  a four-frame random-access GOF at 0.83 Mpts, occupancyPrecision 2, GPA, I=50   (configs[2])
  vox11  one 11-bit frame of ~2.9 Mpts on the 2560-wide canvas, I=20                    (configs[4])
These exercise what the small cases cannot: the refine adjacency capacity retry, hot-set spill in the orientation walk, patch
arenas, packing beyond 1280 rows, push-pull on 2560-wide canvases."""
import importlib.util
import json
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", "golden", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def fullsize_fixture():
    with open(os.path.join(ROOT, "tests", "golden", "gof_fullsize.json")) as f:
        return json.load(f)["cases"]


def digests_of_case(product, name):
    m = _load("make_golden_fullsize")
    frames = m.case_frames(name)
    w = product.weight_normal(frames[0][0], m.case_bits(name))
    prm, prec = m.case_params(name, w)
    return [float(x) for x in w], [len(f[0]) for f in frames], m.products_digest(product.encode_gof(frames, prm, occupancy_precision=prec))


@pytest.mark.parametrize("name", ["ai_r3", "ra_r5", "vox11"])
def test_fullsize_products_equal_the_reference(name, product):
    gold = fullsize_fixture()
    if name not in gold:
        pytest.skip("case %s not in the fixture" % name)
    w, pts, got = digests_of_case(product, name)
    assert pts == gold[name]["points"]
    assert w == gold[name]["weight_normal"]
    bad = ["frame %d %s" % (f, k) for f, (a, b) in enumerate(zip(got, gold[name]["frames"])) for k in b if a[k] != b[k]]
    assert len(got) == len(gold[name]["frames"]) and bad == [], "products differ from the reference at full size: %s" % bad[:8]
