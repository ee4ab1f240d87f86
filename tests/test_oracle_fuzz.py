"""Oracle vs the reference itself on random small clouds of unstructured shapes (noisy blobs, thin noisy shells, crossing noisy
planes, sparse dust): every stage a1-a26 plus the codec-ready YUV frames, all-intra and random access. The shapes are chosen to
produce what the structured fixtures rarely do: frustrated orientation fields, many tiny connected components, ragged patches,
points left unprojected."""
import numpy as np
import pytest

import bindings


def cloud(kind, rng):
    if kind == "blob":          # gaussian blob, voxelised: dense core, dusty rim
        p = rng.normal(60, 9, size=(9000, 3))
    elif kind == "shell":       # noisy sphere shell, radial noise up to +-2
        d = rng.normal(size=(12000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = 64 + d * (28 + rng.uniform(-2, 2, size=(len(d), 1)))
    elif kind == "planes":      # three crossing noisy planes
        a, b = rng.uniform(10, 90, size=(2, 9000))
        n = rng.normal(0, 0.7, 9000)
        k = rng.integers(0, 3, 9000)
        p = np.stack([np.where(k == 0, 50 + n, a), np.where(k == 1, 50 + n, np.where(k == 0, a, b)), np.where(k == 2, 50 + n, b)], 1)
    else:                       # dust: sparse uniform noise plus a small dense cube
        p = np.concatenate([rng.uniform(5, 120, size=(2500, 3)), rng.uniform(40, 56, size=(5000, 3))])
    xyz = np.unique(np.clip(np.rint(p), 0, 1023).astype(np.int16), axis=0)
    xyz = xyz[rng.permutation(len(xyz))]
    rgb = rng.integers(0, 256, size=(len(xyz), 3)).astype(np.uint8)
    rgb[:, 0] = (xyz[:, 0] * 2) % 256     # some structure so that the D1 colour gate sees both outcomes
    return xyz, rgb


@pytest.mark.parametrize("kind", ["blob", "shell", "planes", "dust"])
def test_oracle_vs_reference_on_random_clouds(kind, oracle, reference):
    rng = np.random.default_rng({"blob": 11, "shell": 12, "planes": 13, "dust": 14}[kind])
    frames = [cloud(kind, rng), cloud(kind, rng)]
    for ra, prec in (((0, 4),) if kind in ("blob", "planes") else ((1, 2),)):   # (one packing mode per shape keeps the CPU suite short)
        prm = bindings.ctc_seg_params(bits=10, iterations=6, weight=reference.weight_normal(frames[0][0], 11))
        prm.global_patch_allocation = ra
        ref, _ = reference.encode_gof(frames, prm, occupancy_precision=prec)
        assert bindings.compare_gof(oracle.encode_gof(frames, prm, occupancy_precision=prec), ref) == [], "%s ra=%d" % (kind, ra)
        assert sum(len(f.patches.patches) for f in ref) > 2
