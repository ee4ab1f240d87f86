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
