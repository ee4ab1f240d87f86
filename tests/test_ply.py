"""Input side: pccb200_ply_read against the reference's own PCCPointSet3::read (oracle/_ref) on files that
exercise its grammar: ascii with integer / decimal / exponent tokens, blank lines, comments, CRLF, extra properties; binary with
float32 / float64 / uint16 coordinates and interleaved extra properties; files without colours; corrupt files.  Host code only:
runs on CPU."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import bindings
import synth


def product_lib():
    lib = C.CDLL(bindings.PRODUCT_SO)
    lib.pccb200_ply_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    return lib


def read_with(fn, path):
    n, col = C.c_size_t(0), C.c_int(-1)
    rc = fn(path.encode(), None, None, 0, C.byref(n), C.byref(col))
    if rc != 0:
        return rc, None, None, None
    xyz = np.full((n.value, 3), -7, np.int16)
    rgb = np.full((n.value, 3), 9, np.uint8)
    rc = fn(path.encode(), xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p), n.value, C.byref(n), C.byref(col))
    return rc, xyz, rgb, col.value


def header(n, fmt, props):
    return ("ply\nformat %s 1.0\ncomment made by tests/test_ply.py\nelement vertex %d\n" % (fmt, n) + "".join("property %s %s\n" % p for p in props) +
            "element face 0\nproperty list uchar int vertex_index\nend_header\n").encode()


def write_cases(tmp):
    xyz, rgb = synth.figure(scale=0.2, seed=1, frame=0)
    xyz, rgb = xyz[:60000], rgb[:60000]
    n = len(xyz)
    out = {}
    body = "".join("%d %d %d %d %d %d\n" % (*p, *c) for p, c in zip(xyz.tolist(), rgb.tolist()))
    out["ascii_int"] = header(n, "ascii", [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")]) + body.encode()
    lines = []
    for i, (p, c) in enumerate(zip(xyz.tolist(), rgb.tolist())):   # decimals, exponents, signs, tabs, CRLF, blank lines, extra columns
        if i % 997 == 0:
            lines.append("\r\n")
        tok = ["%.6f" % p[0], "%.3e" % (p[1] + 0.25), "+%d." % p[2], "0.5", "-0.25", "1.0", str(c[0]), "%d" % (c[1] + 256), str(c[2]), "77"]
        lines.append(("\t".join(tok) if i % 3 == 0 else " ".join(tok)) + ("\r\n" if i % 5 == 0 else "\n"))
    out["ascii_mixed"] = header(n, "ascii", [("double", "x"), ("float", "y"), ("float", "z"), ("float", "nx"), ("float", "ny"), ("float", "nz"), ("uchar", "red"),
                                            ("uchar", "green"), ("uchar", "blue"), ("uchar", "alpha")]) + "".join(lines).encode()
    out["ascii_no_colour"] = header(n, "ascii", [("float", "x"), ("float", "y"), ("float", "z")]) + "".join("%d %d %d\n" % tuple(p) for p in xyz.tolist()).encode()
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("refl", "<u2")])
    rec["x"], rec["y"], rec["z"], rec["nx"] = xyz[:, 0] + 0.75, xyz[:, 1], xyz[:, 2] + 0.5, 0.1
    rec["r"], rec["g"], rec["b"], rec["refl"] = rgb[:, 0], rgb[:, 1], rgb[:, 2], 300
    out["binary_f32"] = header(n, "binary_little_endian", [("float", "x"), ("float", "y"), ("float", "z"), ("float", "nx"), ("uchar", "red"), ("uchar", "green"),
                                                            ("uchar", "blue"), ("uint16", "confidence")]) + rec.tobytes()   # (a 2-byte "reflectance" crashes the reference reader: PCCPointSet.cpp:743)
    rec = np.zeros(n, dtype=[("r", "u1"), ("x", "<f8"), ("g", "u1"), ("y", "<u2"), ("b", "u1"), ("z", "<f8")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0] + 0.999, xyz[:, 1], xyz[:, 2]
    rec["r"], rec["g"], rec["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    out["binary_f64_u16"] = header(n, "binary_little_endian", [("uchar", "red"), ("double", "x"), ("uchar", "green"), ("uint16", "y"), ("uchar", "blue"),
                                                                ("double", "z")]) + rec.tobytes()
    out["more_lines_than_points"] = header(100, "ascii", [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")]) + body[:20000].encode()
    paths = {}
    for k, v in out.items():
        paths[k] = os.path.join(tmp, k + ".ply")
        with open(paths[k], "wb") as f:
            f.write(v)
    return paths


def test_ply_read_matches_reference(tmp_path, reference):
    lib = product_lib()
    reference.lib.ref_read_ply.argtypes = lib.pccb200_ply_read.argtypes
    for name, path in write_cases(str(tmp_path)).items():
        r_rc, r_xyz, r_rgb, r_col = read_with(reference.lib.ref_read_ply, path)
        p_rc, p_xyz, p_rgb, p_col = read_with(lib.pccb200_ply_read, path)
        assert r_rc == 0 and p_rc == 0, name
        assert p_col == r_col and p_xyz.shape == r_xyz.shape, name
        assert np.array_equal(p_xyz, r_xyz), "%s positions" % name
        if r_col:
            assert np.array_equal(p_rgb, r_rgb), "%s colours" % name


def test_ply_read_rejects_what_the_reference_rejects(tmp_path, reference):
    lib = product_lib()
    reference.lib.ref_read_ply.argtypes = lib.pccb200_ply_read.argtypes
    bad = {"not_ply": b"plx\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n1\n",
           "no_z": header(1, "ascii", [("float", "x"), ("float", "y")]) + b"1 2\n",
           "version": b"ply\nformat ascii 2.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n",
           "short_line": header(2, "ascii", [("float", "x"), ("float", "y"), ("float", "z")]) + b"1 2 3\n4 5\n"}
    for name, data in bad.items():
        path = os.path.join(str(tmp_path), name + ".ply")
        with open(path, "wb") as f:
            f.write(data)
        assert read_with(reference.lib.ref_read_ply, path)[0] != 0, name
        assert read_with(lib.pccb200_ply_read, path)[0] != 0, name
    assert read_with(lib.pccb200_ply_read, os.path.join(str(tmp_path), "missing.ply"))[0] != 0


def test_ply_read_is_faster_than_the_reference(tmp_path, reference):
    """not a benchmark, a sanity bound: one 0.2 Mpts ascii frame"""
    import time
    lib = product_lib()
    reference.lib.ref_read_ply.argtypes = lib.pccb200_ply_read.argtypes
    xyz, rgb = synth.figure(scale=0.3, seed=2, frame=0)
    path = os.path.join(str(tmp_path), "frame.ply")
    with open(path, "wb") as f:
        f.write(header(len(xyz), "ascii", [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")]))
        f.write("".join("%d %d %d %d %d %d\n" % (*p, *c) for p, c in zip(xyz.tolist(), rgb.tolist())).encode())
    t = time.perf_counter()
    r = read_with(reference.lib.ref_read_ply, path)
    t_ref = time.perf_counter() - t
    t = time.perf_counter()
    p = read_with(lib.pccb200_ply_read, path)
    t_b200 = time.perf_counter() - t
    assert np.array_equal(p[1], xyz) and np.array_equal(p[2], rgb) and np.array_equal(r[1], xyz)
    assert t_b200 < t_ref, (t_b200, t_ref)


# ---- PCCGroupOfFrames::load (PccLibCommon/source/PCCGroupOfFrames.cpp:46-83): pccb200_ply_read_frames + pccb200shim::loadFrames ----------

def write_group(tmp, first, kinds):
    """frames first, first+1, ... as <tmp>/frame_%04d.ply, copies of the single-frame cases named in `kinds`; None leaves a hole"""
    import shutil
    cases = write_cases(tmp)
    for k, kind in enumerate(kinds):
        if kind is not None:
            shutil.copyfile(cases[kind], os.path.join(tmp, "frame_%04d.ply" % (first + k)))
    return os.path.join(tmp, "frame_%04d.ply")


def read_frames(lib, pattern, start, end, threads):
    lib.pccb200_ply_read_frames.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.POINTER(C.c_size_t)]
    count = end - start
    n, col, good = np.zeros(count, np.uint64), np.full(count, -1, np.int32), C.c_size_t(99)
    rc = lib.pccb200_ply_read_frames(pattern.encode(), start, end, None, None, None, n.ctypes.data, col.ctypes.data, threads, C.byref(good))
    xyz = [np.full((int(n[k]), 3), -7, np.int16) for k in range(good.value)]
    rgb = [np.full((int(n[k]), 3), 9, np.uint8) for k in range(good.value)]
    if good.value:
        xp = (C.c_void_p * good.value)(*[a.ctypes.data for a in xyz])
        cp = (C.c_void_p * good.value)(*[a.ctypes.data for a in rgb])
        cap, n2, read = n[:good.value].copy(), np.zeros(good.value, np.uint64), C.c_size_t(99)
        rc2 = lib.pccb200_ply_read_frames(pattern.encode(), start, start + good.value, xp, cp, cap.ctypes.data, n2.ctypes.data, None, threads, C.byref(read))
        assert rc2 == 0 and read.value == good.value and np.array_equal(n2, cap)
    return rc, good.value, xyz, rgb, col


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_ply_read_frames_equals_frame_by_frame_reads(tmp_path, threads):
    lib = product_lib()
    kinds = ["ascii_int", "binary_f32", "ascii_mixed", "ascii_no_colour", "binary_f64_u16"]
    pattern = write_group(str(tmp_path), 1051, kinds)
    rc, good, xyz, rgb, col = read_frames(lib, pattern, 1051, 1051 + len(kinds), threads)
    assert rc == 0 and good == len(kinds)
    for k in range(good):
        one = read_with(lib.pccb200_ply_read, pattern % (1051 + k))
        assert one[0] == 0 and col[k] == one[3] and np.array_equal(xyz[k], one[1]), kinds[k]
        if one[3]:
            assert np.array_equal(rgb[k], one[2]), kinds[k]


def test_ply_read_frames_ends_the_group_at_the_first_unreadable_frame(tmp_path):
    lib = product_lib()
    pattern = write_group(str(tmp_path), 7, ["ascii_int", "binary_f32", None, "ascii_no_colour"])
    rc, good, xyz, rgb, col = read_frames(lib, pattern, 7, 11, 4)
    assert rc != 0 and good == 2 and len(xyz[1]) == 60000
    rc, good, *_ = read_frames(lib, pattern, 9, 11, 4)
    assert rc != 0 and good == 0
    rc, good, *_ = read_frames(lib, pattern, 8, 8, 4)       # an empty range is an empty group
    assert rc == 0 and good == 0
    good = C.c_size_t(5)
    n = np.zeros(1, np.uint64)
    assert lib.pccb200_ply_read_frames(pattern.encode(), 9, 8, None, None, None, n.ctypes.data, None, 1, C.byref(good)) != 0 and good.value == 0


def shim_lib():
    if not os.path.exists(bindings.SHIM_SO):
        pytest.skip("oracle/_ref/libtmc2shim.so not built (needs /root/reference)")
    lib = C.CDLL(bindings.SHIM_SO)
    lib.shim_load_compare.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


@pytest.mark.parametrize("color_transform", [0, 1])
def test_shim_load_frames_fills_the_group_like_the_reference(tmp_path, color_transform):
    """pccb200shim::loadFrames against PCCGroupOfFrames::load inside the reference's own classes (oracle/shim_harness.cpp):
    return value, frame count, point counts, colour flags, positions, colours (after the RGB -> YUV point transform when asked)"""
    lib = shim_lib()
    kinds = ["ascii_int", "binary_f32", "ascii_mixed", "ascii_no_colour", "binary_f64_u16", "more_lines_than_points"]
    pattern = write_group(str(tmp_path), 1, kinds)
    frames, points = C.c_size_t(), C.c_size_t()
    assert lib.shim_load_compare(pattern.encode(), 1, 1 + len(kinds), color_transform, 4, C.byref(frames), C.byref(points), None, None) == 0
    assert frames.value == len(kinds) and points.value == 5 * 60000 + 100
    # a hole ends the group in both; an empty range returns false in both; a short body line ends it in both
    os.remove(pattern % 4)
    assert lib.shim_load_compare(pattern.encode(), 1, 1 + len(kinds), color_transform, 2, C.byref(frames), C.byref(points), None, None) == 0
    assert frames.value == 3
    assert lib.shim_load_compare(pattern.encode(), 4, 6, color_transform, 2, C.byref(frames), C.byref(points), None, None) == 0 and frames.value == 0
    assert lib.shim_load_compare(pattern.encode(), 2, 2, color_transform, 2, C.byref(frames), C.byref(points), None, None) == 0 and frames.value == 0
    with open(pattern % 3, "wb") as f:
        f.write(header(3, "ascii", [("float", "x"), ("float", "y"), ("float", "z")]) + b"1 2 3\n4 5\n6 7 8\n")
    assert lib.shim_load_compare(pattern.encode(), 1, 4, color_transform, 2, C.byref(frames), C.byref(points), None, None) == 0 and frames.value == 2


def test_shim_load_frames_is_faster_than_the_reference(tmp_path):
    """a sanity bound, not a benchmark: four 0.2 Mpts ascii frames"""
    lib = shim_lib()
    for k in range(4):
        xyz, rgb = synth.figure(scale=0.3, seed=2, frame=k)
        with open(os.path.join(str(tmp_path), "f%02d.ply" % k), "wb") as f:
            f.write(header(len(xyz), "ascii", [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")]))
            f.write("".join("%d %d %d %d %d %d\n" % (*p, *c) for p, c in zip(xyz.tolist(), rgb.tolist())).encode())
    t_ref, t_shim, frames, points = C.c_double(), C.c_double(), C.c_size_t(), C.c_size_t()
    pattern = os.path.join(str(tmp_path), "f%02d.ply")
    assert lib.shim_load_compare(pattern.encode(), 0, 4, 0, 8, C.byref(frames), C.byref(points), C.byref(t_ref), C.byref(t_shim)) == 0 and frames.value == 4
    assert t_shim.value < t_ref.value, (t_shim.value, t_ref.value)
