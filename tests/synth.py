"""Deterministic synthetic voxelised point clouds (SURVEY.md §8d S1–S4).

No dataset ships with the reference (cfg/sequence/*.cfg point at external 8i/owlii .ply files), so every
test / bench input is generated here: integer coordinates, no duplicate positions, uint8 colours.
"""
import numpy as np


def _unique_rows(xyz):
    """Deduplicate keeping first occurrences (stable), like a voxelised capture has no duplicate positions."""
    key = (xyz[:, 0].astype(np.int64) << 32) | (xyz[:, 1].astype(np.int64) << 16) | xyz[:, 2].astype(np.int64)
    _, first = np.unique(key, return_index=True)
    first.sort()
    return xyz[first]


def _texture(xyz, seed):
    """Band-limited colour field + noise; exercises the D1 colour-similarity gate."""
    rng = np.random.default_rng(seed + 77)
    f = xyz.astype(np.float64) / 37.0
    r = 128 + 100 * np.sin(f[:, 0] + 0.3 * f[:, 1]) + rng.normal(0, 12, len(xyz))
    g = 128 + 100 * np.sin(1.7 * f[:, 1] - 0.2 * f[:, 2] + 1.0) + rng.normal(0, 12, len(xyz))
    b = 128 + 100 * np.cos(0.9 * f[:, 2] + 0.5 * f[:, 0]) + rng.normal(0, 12, len(xyz))
    return np.clip(np.stack([r, g, b], 1), 0, 255).astype(np.uint8)


def _shell(center, radii, samples, rng, thickness=0.6):
    """Voxelised ellipsoid shell."""
    u = rng.normal(size=(samples, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    p = center + u * radii + rng.uniform(-thickness, thickness, size=(samples, 3))
    return np.rint(p)


def sphere(radius=40, center=128, seed=0, density=6.0):
    """S1: voxelised sphere shell (plumbing case)."""
    rng = np.random.default_rng(seed)
    n = int(density * 4 * np.pi * radius * radius)
    xyz = _shell(np.array([center] * 3, float), np.array([radius] * 3, float), n, rng)
    xyz = _unique_rows(np.clip(xyz, 0, 32767).astype(np.int16))
    return xyz, _texture(xyz, seed)


def figure(scale=1.0, seed=0, bits=10, frame=0):
    """S2: 'longdress-like' union of displaced ellipsoid shells + a folded skirt, voxelised to `bits`.
    scale=1.0 gives ≈0.8 M points at 10 bit. `frame` applies a small rigid + non-rigid motion."""
    rng = np.random.default_rng(seed)
    size = (1 << bits) - 1
    s = scale * size / 1023.0
    t = 0.04 * frame
    cx, cz = 0.5 * size + 6 * s * np.sin(t), 0.5 * size + 4 * s * np.cos(t)
    parts = []
    # (centre offset, radii, relative sample weight)
    body = [((0, 850, 0), (70, 90, 80), 1.0),      # head
            ((0, 640, 0), (120, 150, 85), 2.2),    # torso
            ((-150, 650 + 20 * np.sin(t), 0), (40, 160, 40), 1.0),   # arm
            ((150, 650 - 20 * np.sin(t), 0), (40, 160, 40), 1.0),    # arm
            ((-60, 150, 0), (50, 170, 50), 1.2),   # leg
            ((60, 150, 0), (50, 170, 50), 1.2)]    # leg
    for off, rad, wgt in body:
        area = 4 * np.pi * ((rad[0] * rad[1]) ** 1.6 / 3 + (rad[0] * rad[2]) ** 1.6 / 3 + (rad[1] * rad[2]) ** 1.6 / 3) ** (1 / 1.6)
        n = int(7.0 * area * s * s)
        c = np.array([cx + off[0] * s, off[1] * s, cz + off[2] * s])
        parts.append(_shell(c, np.array(rad, float) * s, n, rng))
    # skirt: cone with sinusoidal folds (thin sheet, both sides visible -> D0/D1 + occlusions)
    n = int(7.0 * 2 * np.pi * 200 * 420 * s * s)
    h = rng.uniform(0, 1, n)
    a = rng.uniform(0, 2 * np.pi, n)
    r = (90 + 210 * h) * (1 + 0.10 * np.sin(9 * a + 3 * h + t) * h)
    sk = np.stack([cx + r * np.cos(a) * s, (560 - 420 * h) * s, cz + r * np.sin(a) * s], 1)
    sk += rng.uniform(-0.6, 0.6, sk.shape)
    parts.append(np.rint(sk))
    xyz = np.concatenate(parts)
    xyz = _unique_rows(np.clip(xyz, 0, size).astype(np.int16))
    return xyz, _texture(xyz, seed + frame)


def planes(n_side=48, seed=0):
    """S4a: axis-aligned planes (an open box): every k-NN distance has maximal ties."""
    g = np.arange(n_side)
    a, b = np.meshgrid(g, g, indexing="ij")
    a, b = a.ravel(), b.ravel()
    z = np.zeros_like(a)
    faces = [np.stack([a, b, z], 1), np.stack([a, z, b], 1), np.stack([z, a, b], 1),
             np.stack([a, b, z + n_side - 1], 1), np.stack([a, z + n_side - 1, b], 1)]
    xyz = _unique_rows((np.concatenate(faces) + 20).astype(np.int16))
    rng = np.random.default_rng(seed)
    xyz = xyz[rng.permutation(len(xyz))]
    return xyz, _texture(xyz, seed)


def double_sheet(n_side=64, gap=3, seed=0):
    """S4b: two parallel sheets `gap` apart (≤ surfaceThickness → D0/D1 pairs) with opposite colours."""
    g = np.arange(n_side)
    a, b = np.meshgrid(g, g, indexing="ij")
    a, b = a.ravel(), b.ravel()
    wob = np.rint(2 * np.sin(a / 9.0) + 2 * np.cos(b / 7.0)).astype(np.int64)
    s0 = np.stack([a + 30, b + 30, 60 + wob], 1)
    s1 = np.stack([a + 30, b + 30, 60 + wob + gap], 1)
    xyz = np.concatenate([s0, s1]).astype(np.int16)
    rgb = np.concatenate([np.tile([[200, 40, 40]], (len(s0), 1)), np.tile([[210, 60, 50]], (len(s1), 1))]).astype(np.uint8)
    rng = np.random.default_rng(seed)
    flip = rng.random(len(s1)) < 0.2          # 20 % of the back sheet fails the colour gate
    rgb[len(s0):][flip] = [10, 250, 10]
    perm = rng.permutation(len(xyz))
    return xyz[perm], rgb[perm]


def specks(seed=0):
    """S4c: a small sphere plus isolated specks of <16 points (dropped connected components)."""
    xyz, rgb = sphere(radius=18, center=64, seed=seed)
    rng = np.random.default_rng(seed + 5)
    extra = []
    for _ in range(12):
        c = rng.integers(100, 200, size=3)
        k = rng.integers(1, 15)
        extra.append(c + rng.integers(-1, 2, size=(k, 3)))
    e = _unique_rows(np.concatenate(extra).astype(np.int16))
    xyz2 = _unique_rows(np.concatenate([xyz, e]))
    return xyz2, _texture(xyz2, seed)


def random_cloud(n=2000, span=64, seed=0):
    """Uniform random unique integer points (kd-tree / k-NN stress: no structure, many ties)."""
    rng = np.random.default_rng(seed)
    xyz = _unique_rows(rng.integers(0, span, size=(n * 2, 3)).astype(np.int16))[:n]
    return xyz, _texture(xyz, seed)


def sheet_stack(layers=14, frame=0, seed=0, spacing=3, extent=126, gap=14):
    """Packing stress (random-access / global patch allocation): 9 x `layers` sparse square sheets, each its own connected
    component and an (extent/16)^2-block patch, jittered and resized from frame to frame so that patch unions grow and the
    packed canvas exceeds the minimum image height. Few points per occupancy block keep segmentation cheap."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=(3, 3, layers, 2))          # per-sheet origin jitter (fixed over time)
    grow = rng.integers(0, 3, size=(3, 3, layers))
    parts = []
    for ix in range(3):
        for iy in range(3):
            for iz in range(layers):
                e = extent + 16 * int((grow[ix, iy, iz] + frame * (1 + (ix + iy + iz) % 2)) % 3) - 16
                g = np.arange(0, e, spacing)
                a, b = np.meshgrid(g, g, indexing="ij")
                a, b = a.ravel(), b.ravel()
                ox = 8 + ix * 170 + int(base[ix, iy, iz, 0]) + (frame * (1 + iz % 3)) % 5
                oy = 8 + iy * 170 + int(base[ix, iy, iz, 1]) + (frame * (1 + ix % 2)) % 4
                oz = 6 + iz * gap
                parts.append(np.stack([a + ox, b + oy, np.full_like(a, oz)], 1))
    xyz = _unique_rows(np.concatenate(parts).astype(np.int16))
    xyz = xyz[np.random.default_rng(seed + frame).permutation(len(xyz))]
    return xyz, _texture(xyz, seed + frame)
