"""ctypes bindings used by the tests: the CPU oracle (oracle/libpccoracle.so), the compiled reference
(oracle/_ref/libtmc2ref.so, optional: present only where it was built) and the product (libpccb200.so).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libpccoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtmc2ref.so")
PRODUCT_SO = os.path.join(ROOT, "mpeg-pcc-tmc2_b200", "libpccb200.so")

c_i16p = C.POINTER(C.c_int16)
c_u8p = C.POINTER(C.c_uint8)
c_u32p = C.POINTER(C.c_uint32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u64p = C.POINTER(C.c_uint64)


class SegParams(C.Structure):
    """mirror of pccb200_seg_params (include/pccb200.h)"""
    _fields_ = [(n, C.c_int32) for n in (
        "nn_normal_estimation", "normal_orientation", "max_nn_count_refine", "iteration_count_refine",
        "voxel_dim_refine", "search_radius_refine", "occupancy_resolution", "enable_patch_splitting",
        "max_patch_size", "quantizer_size_x", "quantizer_size_y", "min_point_count_per_cc",
        "max_nn_count_patch_seg", "surface_thickness", "min_level", "max_allowed_depth",
        "geometry_bitdepth_2d", "geometry_bitdepth_3d", "map_count_minus1", "global_patch_allocation")] + [
        ("lambda_refine", C.c_double), ("max_allowed_dist2_raw_detection", C.c_double),
        ("max_allowed_dist2_raw_selection", C.c_double), ("weight_normal", C.c_double * 3)]


def ctc_seg_params(bits=10, iterations=50, weight=(1.0, 1.0, 1.0)):
    """CTC all-intra values."""
    p = SegParams()
    p.nn_normal_estimation = 16
    p.normal_orientation = 1
    p.max_nn_count_refine = 1024
    p.iteration_count_refine = iterations
    p.voxel_dim_refine = 4
    p.search_radius_refine = 192
    p.occupancy_resolution = 16
    p.enable_patch_splitting = 1
    p.max_patch_size = 1024
    p.quantizer_size_x = 16
    p.quantizer_size_y = 16
    p.min_point_count_per_cc = 16
    p.max_nn_count_patch_seg = 16
    p.surface_thickness = 4
    p.min_level = 64
    p.max_allowed_depth = 255
    p.geometry_bitdepth_2d = 8
    p.geometry_bitdepth_3d = bits + 1
    p.map_count_minus1 = 1
    p.lambda_refine = 3.0
    p.max_allowed_dist2_raw_detection = 9.0
    p.max_allowed_dist2_raw_selection = 1.0
    for i in range(3):
        p.weight_normal[i] = weight[i]
    return p


class Patch(C.Structure):
    """mirror of pccb200_patch"""
    _fields_ = [(n, C.c_int32) for n in (
        "index", "view_id", "normal_axis", "tangent_axis", "bitangent_axis", "projection_mode",
        "u1", "v1", "d1", "size_u", "size_v", "size_d", "size_d_pixel", "size_u0", "size_v0",
        "size_2d_x", "size_2d_y", "u0", "v0", "orientation", "d0_count", "eom_and_d1_count", "best_match_idx", "is_global")] + [
        ("depth_offset", C.c_int64), ("occ_offset", C.c_int64)]


PATCH_DTYPE = np.dtype([(n, np.int32) for n, _ in Patch._fields_[:24]] + [("depth_offset", np.int64), ("occ_offset", np.int64)])
assert PATCH_DTYPE.itemsize == C.sizeof(Patch)


def ptr(a, t):
    return a.ctypes.data_as(t)


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.int16)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


class PatchSet:
    """patch metadata (structured array) + depth arena (int16) + occupancy arena (uint8)"""

    def __init__(self, patches, depth, occ):
        self.patches, self.depth, self.occ = patches, depth, occ

    def depth_maps(self, i):
        p = self.patches[i]
        px = int(p["size_u"]) * int(p["size_v"])
        o = int(p["depth_offset"])
        shape = (int(p["size_v"]), int(p["size_u"]))
        return self.depth[o:o + px].reshape(shape), self.depth[o + px:o + 2 * px].reshape(shape)

    def occupancy(self, i):
        p = self.patches[i]
        nb = int(p["size_u0"]) * int(p["size_v0"])
        o = int(p["occ_offset"])
        return self.occ[o:o + nb].reshape(int(p["size_v0"]), int(p["size_u0"]))


def _collect_patches(lib, prefix, h):
    n = getattr(lib, prefix + "patches_count")(h)
    de = getattr(lib, prefix + "patches_depth_elems")(h)
    oe = getattr(lib, prefix + "patches_occ_elems")(h)
    patches = np.zeros(n, dtype=PATCH_DTYPE)
    depth = np.zeros(de, dtype=np.int16)
    occ = np.zeros(oe, dtype=np.uint8)
    getattr(lib, prefix + "patches_get")(h, patches.ctypes.data_as(C.c_void_p), ptr(depth, c_i16p), ptr(occ, c_u8p))
    getattr(lib, prefix + "patches_free")(h)
    return PatchSet(patches, depth, occ)


def _decl_patch_api(lib, prefix):
    getattr(lib, prefix + "patches_count").restype = C.c_int
    getattr(lib, prefix + "patches_count").argtypes = [C.c_void_p]
    for nme in ("patches_depth_elems", "patches_occ_elems"):
        getattr(lib, prefix + nme).restype = C.c_size_t
        getattr(lib, prefix + nme).argtypes = [C.c_void_p]
    getattr(lib, prefix + "patches_get").restype = None
    getattr(lib, prefix + "patches_get").argtypes = [C.c_void_p, C.c_void_p, c_i16p, c_u8p]
    getattr(lib, prefix + "patches_free").restype = None
    getattr(lib, prefix + "patches_free").argtypes = [C.c_void_p]


# ------------------------------------------------------------------------------------------------ oracle
class _Lazy:
    """CDLL wrapper: declaring the signature of a symbol that is not built yet is not an error until it is called."""

    class _Missing:
        def __init__(self, name):
            self.name = name

        def __call__(self, *a):
            raise RuntimeError("symbol %s is not exported by the library" % self.name)

    def __init__(self, path):
        object.__setattr__(self, "_dll", C.CDLL(path))

    def __getattr__(self, name):
        try:
            return getattr(self._dll, name)
        except AttributeError:
            m = _Lazy._Missing(name)
            object.__setattr__(self, name, m)
            return m


class Oracle:
    def __init__(self, path=ORACLE_SO):
        L = self.lib = _Lazy(path)
        L.pcco_kdtree_build.restype = C.c_void_p
        L.pcco_kdtree_build.argtypes = [c_i16p, C.c_size_t]
        L.pcco_kdtree_free.argtypes = [C.c_void_p]
        L.pcco_kdtree_vind.argtypes = [C.c_void_p, c_u32p]
        L.pcco_knn.argtypes = [C.c_void_p, c_i16p, C.c_size_t, C.c_int, c_u32p, c_f32p]
        L.pcco_radius.restype = C.c_size_t
        L.pcco_radius.argtypes = [C.c_void_p, c_i16p, C.c_size_t, C.c_double, C.c_size_t, c_u64p, c_u32p, c_f32p]
        L.pcco_normals.argtypes = [c_i16p, C.c_size_t, c_u32p, C.c_int, c_f64p]
        L.pcco_orient_normals.argtypes = [c_i16p, C.c_size_t, c_u32p, C.c_int, c_f64p]
        L.pcco_weight_normal.argtypes = [c_i16p, C.c_size_t, C.c_int, C.c_double, c_f64p]
        L.pcco_initial_segmentation.argtypes = [c_f64p, C.c_size_t, c_f64p, c_u8p]
        L.pcco_refine_segmentation.argtypes = [c_i16p, c_f64p, C.c_size_t, C.POINTER(SegParams), c_u8p]
        L.pcco_segment_patches.restype = C.c_void_p
        L.pcco_segment_patches.argtypes = [c_i16p, c_u8p, C.c_size_t, c_u32p, C.c_int, c_u8p, C.POINTER(SegParams)]
        _decl_patch_api(L, "pcco_")

    def vind(self, xyz):
        xyz = _xyz(xyz)
        t = self.lib.pcco_kdtree_build(ptr(xyz, c_i16p), len(xyz))
        v = np.zeros(len(xyz), np.uint32)
        self.lib.pcco_kdtree_vind(t, ptr(v, c_u32p))
        self.lib.pcco_kdtree_free(t)
        return v

    def knn(self, xyz, q, k):
        xyz, q = _xyz(xyz), _xyz(q)
        t = self.lib.pcco_kdtree_build(ptr(xyz, c_i16p), len(xyz))
        idx = np.zeros((len(q), k), np.uint32)
        d = np.zeros((len(q), k), np.float32)
        self.lib.pcco_knn(t, ptr(q, c_i16p), len(q), k, ptr(idx, c_u32p), ptr(d, c_f32p))
        self.lib.pcco_kdtree_free(t)
        return idx, d

    def radius(self, xyz, q, r2, max_results):
        xyz, q = _xyz(xyz), _xyz(q)
        t = self.lib.pcco_kdtree_build(ptr(xyz, c_i16p), len(xyz))
        off = np.zeros(len(q) + 1, np.uint64)
        total = self.lib.pcco_radius(t, ptr(q, c_i16p), len(q), r2, max_results, ptr(off, c_u64p), None, None)
        idx = np.zeros(total, np.uint32)
        d = np.zeros(total, np.float32)
        self.lib.pcco_radius(t, ptr(q, c_i16p), len(q), r2, max_results, ptr(off, c_u64p), ptr(idx, c_u32p), ptr(d, c_f32p))
        self.lib.pcco_kdtree_free(t)
        return off, idx, d

    def normals(self, xyz, nbr, orient=True):
        xyz = _xyz(xyz)
        nbr = np.ascontiguousarray(nbr, np.uint32)
        out = np.zeros((len(xyz), 3), np.float64)
        self.lib.pcco_normals(ptr(xyz, c_i16p), len(xyz), ptr(nbr, c_u32p), nbr.shape[1], ptr(out, c_f64p))
        if orient:
            self.lib.pcco_orient_normals(ptr(xyz, c_i16p), len(xyz), ptr(nbr, c_u32p), nbr.shape[1], ptr(out, c_f64p))
        return out

    def weight_normal(self, xyz, bits, min_w=0.6):
        xyz = _xyz(xyz)
        w = np.zeros(3, np.float64)
        self.lib.pcco_weight_normal(ptr(xyz, c_i16p), len(xyz), bits, min_w, ptr(w, c_f64p))
        return w

    def initial_segmentation(self, normals, w):
        normals = np.ascontiguousarray(normals, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        part = np.zeros(len(normals), np.uint8)
        self.lib.pcco_initial_segmentation(ptr(normals, c_f64p), len(normals), ptr(w, c_f64p), ptr(part, c_u8p))
        return part

    def refine_segmentation(self, xyz, normals, partition, params):
        xyz = _xyz(xyz)
        normals = np.ascontiguousarray(normals, np.float64)
        part = np.ascontiguousarray(partition, np.uint8).copy()
        self.lib.pcco_refine_segmentation(ptr(xyz, c_i16p), ptr(normals, c_f64p), len(xyz), C.byref(params), ptr(part, c_u8p))
        return part

    def segment_patches(self, xyz, rgb, nbr, partition, params):
        xyz = _xyz(xyz)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        nbr = np.ascontiguousarray(nbr, np.uint32)
        part = np.ascontiguousarray(partition, np.uint8)
        h = self.lib.pcco_segment_patches(ptr(xyz, c_i16p), ptr(rgb, c_u8p), len(xyz), ptr(nbr, c_u32p), nbr.shape[1],
                                          ptr(part, c_u8p), C.byref(params))
        return _collect_patches(self.lib, "pcco_", h)


    def segment_frame_patches(self, xyz, rgb, params):
        """a1..a11 of one frame: the patch list in creation order (PatchSet)"""
        xyz = _xyz(xyz)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.lib.pcco_segment_frame.restype = C.c_void_p
        self.lib.pcco_segment_frame.argtypes = [c_i16p, c_u8p, C.c_size_t, C.POINTER(SegParams)]
        return _collect_patches(self.lib, "pcco_", self.lib.pcco_segment_frame(ptr(xyz, c_i16p), ptr(rgb, c_u8p), len(xyz), C.byref(params)))


# --------------------------------------------------------------------------------------------- reference
class Reference:
    """The reference itself (compiled from /root/reference by oracle/Makefile)."""

    def __init__(self, path=REF_SO):
        L = self.lib = C.CDLL(path)
        L.ref_knn.argtypes = [c_i16p, C.c_size_t, c_i16p, C.c_size_t, C.c_int, c_u32p, c_f32p]
        L.ref_radius.restype = C.c_size_t
        L.ref_radius.argtypes = [c_i16p, C.c_size_t, c_i16p, C.c_size_t, C.c_double, C.c_size_t, c_u64p, c_u32p, c_f32p]
        L.ref_normals.argtypes = [c_i16p, C.c_size_t, C.c_int, C.c_int, c_f64p]
        L.ref_weight_normal.argtypes = [c_i16p, C.c_size_t, C.c_int, C.c_double, c_f64p]
        L.ref_segment_frame.restype = C.c_void_p
        L.ref_segment_frame.argtypes = [c_i16p, c_u8p, C.c_size_t, C.POINTER(SegParams), c_f64p, c_u8p, c_u8p, c_f64p]
        _decl_patch_api(L, "ref_")

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def knn(self, xyz, q, k):
        xyz, q = _xyz(xyz), _xyz(q)
        idx = np.zeros((len(q), k), np.uint32)
        d = np.zeros((len(q), k), np.float32)
        self.lib.ref_knn(ptr(xyz, c_i16p), len(xyz), ptr(q, c_i16p), len(q), k, ptr(idx, c_u32p), ptr(d, c_f32p))
        return idx, d

    def radius(self, xyz, q, r2, max_results):
        xyz, q = _xyz(xyz), _xyz(q)
        off = np.zeros(len(q) + 1, np.uint64)
        total = self.lib.ref_radius(ptr(xyz, c_i16p), len(xyz), ptr(q, c_i16p), len(q), r2, max_results, ptr(off, c_u64p), None, None)
        idx = np.zeros(total, np.uint32)
        d = np.zeros(total, np.float32)
        self.lib.ref_radius(ptr(xyz, c_i16p), len(xyz), ptr(q, c_i16p), len(q), r2, max_results, ptr(off, c_u64p), ptr(idx, c_u32p), ptr(d, c_f32p))
        return off, idx, d

    def normals(self, xyz, k=16, orient=True):
        xyz = _xyz(xyz)
        out = np.zeros((len(xyz), 3), np.float64)
        self.lib.ref_normals(ptr(xyz, c_i16p), len(xyz), k, 1 if orient else 0, ptr(out, c_f64p))
        return out

    def weight_normal(self, xyz, bits, min_w=0.6):
        xyz = _xyz(xyz)
        w = np.zeros(3, np.float64)
        self.lib.ref_weight_normal(ptr(xyz, c_i16p), len(xyz), bits, min_w, ptr(w, c_f64p))
        return w

    def segment_frame(self, xyz, rgb, params):
        """returns dict(normals, partition0, partition1, patches: PatchSet, seconds)"""
        xyz = _xyz(xyz)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        n = len(xyz)
        normals = np.zeros((n, 3), np.float64)
        p0 = np.zeros(n, np.uint8)
        p1 = np.zeros(n, np.uint8)
        sec = C.c_double(0)
        h = self.lib.ref_segment_frame(ptr(xyz, c_i16p), ptr(rgb, c_u8p), n, C.byref(params), ptr(normals, c_f64p),
                                       ptr(p0, c_u8p), ptr(p1, c_u8p), C.byref(sec))
        return dict(normals=normals, partition0=p0, partition1=p1, patches=_collect_patches(self.lib, "ref_", h),
                    seconds=sec.value)


# ----------------------------------------------------------------------------------------------- product
class Product:
    """libpccb200.so through its C ABI (include/pccb200.h). Needs a CUDA device; there is no CPU path."""

    def __init__(self, device=0, path=PRODUCT_SO):
        if not os.path.exists(path):
            raise RuntimeError("libpccb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = self.lib = C.CDLL(path)
        L.pccb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.pccb200_destroy.argtypes = [C.c_void_p]
        L.pccb200_last_error.restype = C.c_char_p
        L.pccb200_last_error.argtypes = [C.c_void_p]
        L.pccb200_version.restype = C.c_char_p
        L.pccb200_knn.argtypes = [C.c_void_p, c_i16p, C.c_size_t, c_i16p, C.c_size_t, C.c_int, c_u32p, c_f32p]
        L.pccb200_kdtree_order.argtypes = [C.c_void_p, c_i16p, C.c_size_t, c_u32p]
        L.pccb200_normals.argtypes = [C.c_void_p, c_i16p, C.c_size_t, C.c_int, C.c_int, c_f64p]
        L.pccb200_segment_frame.argtypes = [C.c_void_p, c_i16p, c_u8p, C.c_size_t, C.POINTER(SegParams), c_f64p, c_u8p, c_u8p,
                                            C.POINTER(C.c_void_p)]
        L.pccb200_patches_count.argtypes = [C.c_void_p]
        L.pccb200_patches_depth_elems.restype = C.c_size_t
        L.pccb200_patches_depth_elems.argtypes = [C.c_void_p]
        L.pccb200_patches_occ_elems.restype = C.c_size_t
        L.pccb200_patches_occ_elems.argtypes = [C.c_void_p]
        L.pccb200_patches_get.argtypes = [C.c_void_p, C.c_void_p, c_i16p, c_u8p]
        L.pccb200_patches_free.argtypes = [C.c_void_p]
        L.pccb200_weight_normal.argtypes = [C.c_void_p, c_i16p, C.c_size_t, C.c_int, C.c_double, c_f64p]
        L.pccb200_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.pccb200_profile_read.argtypes = [C.c_void_p, C.c_char_p, c_f32p, c_f32p, C.c_int, C.POINTER(C.c_int)]
        self.ctx = C.c_void_p()
        rc = L.pccb200_create(device, C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError("pccb200_create failed with %d (no CUDA device? the product has no CPU fallback)" % rc)

    def close(self):
        if self.ctx:
            self.lib.pccb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("pccb200 error %d: %s" % (rc, self.lib.pccb200_last_error(self.ctx).decode()))

    def knn(self, xyz, q, k):
        xyz = _xyz(xyz)
        nq = len(xyz) if q is None else len(q)
        idx = np.zeros((nq, k), np.uint32)
        d = np.zeros((nq, k), np.float32)
        qp = None if q is None else ptr(_xyz(q), c_i16p)
        self._check(self.lib.pccb200_knn(self.ctx, ptr(xyz, c_i16p), len(xyz), qp, nq, k, ptr(idx, c_u32p), ptr(d, c_f32p)))
        return idx, d

    def vind(self, xyz):
        xyz = _xyz(xyz)
        v = np.zeros(len(xyz), np.uint32)
        self._check(self.lib.pccb200_kdtree_order(self.ctx, ptr(xyz, c_i16p), len(xyz), ptr(v, c_u32p)))
        return v

    def normals(self, xyz, k=16, orient=False):
        xyz = _xyz(xyz)
        out = np.zeros((len(xyz), 3), np.float64)
        self._check(self.lib.pccb200_normals(self.ctx, ptr(xyz, c_i16p), len(xyz), k, 1 if orient else 0, ptr(out, c_f64p)))
        return out

    def segment_frame(self, xyz, rgb, params):
        """returns dict(normals, partition0, partition1, patches: PatchSet) like Reference.segment_frame"""
        xyz = _xyz(xyz)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        n = len(xyz)
        normals = np.zeros((n, 3), np.float64)
        p0 = np.zeros(n, np.uint8)
        p1 = np.zeros(n, np.uint8)
        h = C.c_void_p()
        self._check(self.lib.pccb200_segment_frame(self.ctx, ptr(xyz, c_i16p), ptr(rgb, c_u8p), n, C.byref(params),
                                                   ptr(normals, c_f64p), ptr(p0, c_u8p), ptr(p1, c_u8p), C.byref(h)))
        L = self.lib
        cnt = L.pccb200_patches_count(h)
        patches = np.zeros(cnt, dtype=PATCH_DTYPE)
        depth = np.zeros(L.pccb200_patches_depth_elems(h), np.int16)
        occ = np.zeros(L.pccb200_patches_occ_elems(h), np.uint8)
        self._check(L.pccb200_patches_get(h, patches.ctypes.data_as(C.c_void_p), ptr(depth, c_i16p), ptr(occ, c_u8p)))
        L.pccb200_patches_free(h)
        return dict(normals=normals, partition0=p0, partition1=p1, patches=PatchSet(patches, depth, occ))

    def profile(self, on=True):
        self._check(self.lib.pccb200_profile_enable(self.ctx, 1 if on else 0))

    def profile_read(self):
        """[(stage name, device milliseconds)] recorded since the last read"""
        cap = 4096
        names = C.create_string_buffer(32 * cap)
        ms = np.zeros(cap, np.float32)
        st = np.zeros(cap, np.float32)
        cnt = C.c_int(0)
        self._check(self.lib.pccb200_profile_read(self.ctx, names, ptr(ms, c_f32p), ptr(st, c_f32p), cap, C.byref(cnt)))
        out = []
        for i in range(min(cnt.value, cap)):
            out.append((names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), float(ms[i]), float(st[i])))
        return out

    def weight_normal(self, xyz, bits, min_w=0.6):
        xyz = _xyz(xyz)
        w = np.zeros(3, np.float64)
        self._check(self.lib.pccb200_weight_normal(self.ctx, ptr(xyz, c_i16p), len(xyz), bits, min_w, ptr(w, c_f64p)))
        return w


# ---------------------------------------------------------------------------------------- GOF-level products
GOF_OCCUPANCY, GOF_OM_VIDEO, GOF_BLOCK_TO_PATCH, GOF_GEO0, GOF_GEO1, GOF_REC_XYZ, GOF_POINT_TO_PIXEL = 1, 2, 3, 4, 5, 6, 7
GOF_REC_PARTITION, GOF_REC_BOUNDARY, GOF_REC_RGB, GOF_ATTR0_RAW, GOF_ATTR1_RAW, GOF_ATTR0, GOF_ATTR1 = 8, 9, 10, 11, 12, 13, 14
GOF_ATTR0_YUV420, GOF_ATTR1_YUV420 = 15, 16
GOF_GEO0_LUMA8, GOF_GEO1_LUMA8 = 17, 18   # product-only hand-off forms (not produced by the oracle / reference harness: == GEO0 / GEO1 as bytes)
GOF_EXTRA_DTYPES = {17: np.uint8, 18: np.uint8}
GOF_DTYPES = {1: np.uint8, 2: np.uint8, 3: np.uint32, 4: np.uint16, 5: np.uint16, 6: np.int16, 7: np.uint32, 8: np.uint32,
              9: np.uint16, 10: np.uint8, 11: np.uint16, 12: np.uint16, 13: np.uint16, 14: np.uint16, 15: np.uint8, 16: np.uint8}


class GofFrame:
    """products of one frame: dict-like access by GOF_* id, plus packed patches and canvas size"""

    def __init__(self):
        self.patches, self.width, self.height, self.data = None, 0, 0, {}

    def __getitem__(self, what):
        return self.data[what]


def _frames_args(frames):
    n = len(frames)
    xs = [np.ascontiguousarray(f[0], np.int16) for f in frames]
    cs = [np.ascontiguousarray(f[1], np.uint8) for f in frames]
    xp = (c_i16p * n)(*[ptr(x, c_i16p) for x in xs])
    cp = (c_u8p * n)(*[ptr(c, c_u8p) for c in cs])
    ns = (C.c_size_t * n)(*[len(x) for x in xs])
    return n, xs, cs, xp, cp, ns


def _collect_patches_borrowed(lib, prefix, h):
    n = getattr(lib, prefix + "patches_count")(h)
    patches = np.zeros(n, dtype=PATCH_DTYPE)
    depth = np.zeros(getattr(lib, prefix + "patches_depth_elems")(h), dtype=np.int16)
    occ = np.zeros(getattr(lib, prefix + "patches_occ_elems")(h), dtype=np.uint8)
    getattr(lib, prefix + "patches_get")(h, patches.ctypes.data_as(C.c_void_p), ptr(depth, c_i16p), ptr(occ, c_u8p))
    return PatchSet(patches, depth, occ)


def _ref_encode_gof(self, frames, params, occupancy_precision=4, stop_after=0):
    """runs the reference's own stages over a GOF; returns (list of GofFrame, seconds[8])"""
    L = self.lib
    L.ref_encode_gof.restype = C.c_void_p
    L.ref_encode_gof.argtypes = [C.c_int, C.POINTER(c_i16p), C.POINTER(c_u8p), C.POINTER(C.c_size_t), C.POINTER(SegParams), C.c_int, C.c_int]
    L.ref_gof_free.argtypes = [C.c_void_p]
    L.ref_gof_seconds.argtypes = [C.c_void_p, c_f64p]
    L.ref_gof_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.ref_gof_patches.restype = C.c_void_p
    L.ref_gof_patches.argtypes = [C.c_void_p, C.c_int]
    L.ref_gof_get.restype = C.c_size_t
    L.ref_gof_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    n, xs, cs, xp, cp, ns = _frames_args(frames)
    h = L.ref_encode_gof(n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after)
    out = []
    for f in range(n):
        g = GofFrame()
        w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
        L.ref_gof_dims(h, f, C.byref(w), C.byref(hh), C.byref(r))
        g.width, g.height = w.value, hh.value
        g.patches = _collect_patches_borrowed(L, "ref_", L.ref_gof_patches(h, f))
        for what, dt in GOF_DTYPES.items():
            cnt = L.ref_gof_get(h, f, what, None)
            a = np.zeros(cnt, dt)
            if cnt:
                L.ref_gof_get(h, f, what, a.ctypes.data_as(C.c_void_p))
            g.data[what] = a
        out.append(g)
    sec = np.zeros(8)
    L.ref_gof_seconds(h, ptr(sec, c_f64p))
    L.ref_gof_free(h)
    return out, sec


Reference.encode_gof = _ref_encode_gof

SHIM_SO = os.path.join(os.path.dirname(REF_SO), "libtmc2shim.so")


class Shim:
    """the reference's encoder harness with its hot path replaced by the reference-side shim of libpccb200
    (integration/pccb200_shim.cpp; oracle/Makefile target `shim`). Products are read from the reference's own data structures."""

    def __init__(self, path=SHIM_SO):
        self.ref = Reference()
        self.lib = C.CDLL(path)
        self.lib.shim_encode_gof.restype = C.c_void_p
        self.lib.shim_encode_gof.argtypes = [C.c_int, C.POINTER(c_i16p), C.POINTER(c_u8p), C.POINTER(C.c_size_t), C.POINTER(SegParams), C.c_int,
                                             C.c_int, C.POINTER(C.c_int)]

    @staticmethod
    def available():
        return os.path.exists(SHIM_SO) and os.path.exists(REF_SO) and os.path.exists(PRODUCT_SO)

    def decode_gof(self, frames, params, occupancy_precision=4):
        """the reference's stages up to the geometry images, then every frame reconstructed through the decoder-side binding
        (pccb200shim::decodeFrame); products up to generatePointCloud (stop_after = 3)"""
        return self.encode_gof(frames, params, occupancy_precision, 3, entry="shim_decode_gof")

    def encode_gof(self, frames, params, occupancy_precision=4, stop_after=0, entry="shim_encode_gof"):
        """returns (list of GofFrame, status code of the last pccb200 call; frames are empty when the hot path failed)"""
        L = self.ref.lib
        L.ref_gof_free.argtypes = [C.c_void_p]
        L.ref_gof_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.ref_gof_patches.restype = C.c_void_p
        L.ref_gof_patches.argtypes = [C.c_void_p, C.c_int]
        L.ref_gof_get.restype = C.c_size_t
        L.ref_gof_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        n, xs, cs, xp, cp, ns = _frames_args(frames)
        code = C.c_int(0)
        fn = getattr(self.lib, entry)
        fn.restype, fn.argtypes = self.lib.shim_encode_gof.restype, self.lib.shim_encode_gof.argtypes
        h = fn(n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after, C.byref(code))
        out = []
        if code.value == 0:
            for f in range(n):
                g = GofFrame()
                w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
                L.ref_gof_dims(h, f, C.byref(w), C.byref(hh), C.byref(r))
                g.width, g.height = w.value, hh.value
                g.patches = _collect_patches_borrowed(L, "ref_", L.ref_gof_patches(h, f))
                for what, dt in GOF_DTYPES.items():
                    cnt = L.ref_gof_get(h, f, what, None)
                    a = np.zeros(cnt, dt)
                    if cnt:
                        L.ref_gof_get(h, f, what, a.ctypes.data_as(C.c_void_p))
                    g.data[what] = a
                out.append(g)
        L.ref_gof_free(h)
        return out, code.value


def _generic_encode_gof(lib, prefix, frames, params, occupancy_precision=4, stop_after=0, canvas=None):
    g = lambda name: getattr(lib, prefix + name)
    g("encode_gof").restype = C.c_void_p
    g("encode_gof").argtypes = [C.c_int, C.POINTER(c_i16p), C.POINTER(c_u8p), C.POINTER(C.c_size_t), C.POINTER(SegParams), C.c_int, C.c_int]
    if canvas is not None:
        g("encode_gof_canvas").restype = C.c_void_p
        g("encode_gof_canvas").argtypes = g("encode_gof").argtypes + [C.c_size_t, C.c_size_t]
    g("gof_free").argtypes = [C.c_void_p]
    g("gof_dims").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    g("gof_patches").restype = C.c_void_p
    g("gof_patches").argtypes = [C.c_void_p, C.c_int]
    g("gof_get").restype = C.c_size_t
    g("gof_get").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    n, xs, cs, xp, cp, ns = _frames_args(frames)
    if canvas is not None:
        h = g("encode_gof_canvas")(n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after, canvas[0], canvas[1])
    else:
        h = g("encode_gof")(n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after)
    out = []
    for f in range(n):
        fr = GofFrame()
        w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
        g("gof_dims")(h, f, C.byref(w), C.byref(hh), C.byref(r))
        fr.width, fr.height = w.value, hh.value
        fr.patches = _collect_patches_borrowed(lib, prefix, g("gof_patches")(h, f))
        for what, dt in GOF_DTYPES.items():
            cnt = g("gof_get")(h, f, what, None)
            a = np.zeros(cnt, dt)
            if cnt:
                g("gof_get")(h, f, what, a.ctypes.data_as(C.c_void_p))
            fr.data[what] = a
        out.append(fr)
    g("gof_free")(h)
    return out


def _oracle_encode_gof(self, frames, params, occupancy_precision=4, stop_after=0, canvas=None):
    return _generic_encode_gof(self.lib._dll, "pcco_", frames, params, occupancy_precision, stop_after, canvas)


Oracle.encode_gof = _oracle_encode_gof
GOF_NAMES = {1: "occupancy", 2: "om_video", 3: "block_to_patch", 4: "geo0", 5: "geo1", 6: "rec_xyz", 7: "point_to_pixel", 8: "rec_partition",
             9: "rec_boundary", 10: "rec_rgb", 11: "attr0_raw", 12: "attr1_raw", 13: "attr0", 14: "attr1", 15: "attr0_yuv420",
             16: "attr1_yuv420"}


def compare_gof(got, want, attr_tol=0):
    """returns a list of mismatch descriptions (empty == parity). attribute planes / colours may differ by <= attr_tol."""
    bad = []
    if len(got) != len(want):
        return ["frame count %d != %d" % (len(got), len(want))]
    for f, (a, b) in enumerate(zip(got, want)):
        if (a.width, a.height) != (b.width, b.height):
            bad.append("frame %d canvas %dx%d != %dx%d" % (f, a.width, a.height, b.width, b.height))
        if len(a.patches.patches) != len(b.patches.patches):
            bad.append("frame %d patch count %d != %d" % (f, len(a.patches.patches), len(b.patches.patches)))
        else:
            for fld in a.patches.patches.dtype.names:
                if not np.array_equal(a.patches.patches[fld], b.patches.patches[fld]):
                    bad.append("frame %d patch field %s" % (f, fld))
            if not np.array_equal(a.patches.depth, b.patches.depth):
                bad.append("frame %d patch depth arena" % f)
            if not np.array_equal(a.patches.occ, b.patches.occ):
                bad.append("frame %d patch occupancy arena" % f)
        for what, name in GOF_NAMES.items():
            x, y = a.data[what], b.data[what]
            if x.shape != y.shape:
                bad.append("frame %d %s size %d != %d" % (f, name, x.size, y.size))
            elif what in (10, 11, 12, 13, 14, 15, 16) and attr_tol > 0:
                d = np.abs(x.astype(np.int32) - y.astype(np.int32))
                if d.size and d.max() > attr_tol:
                    bad.append("frame %d %s max abs diff %d (> %d) at %d samples" % (f, name, d.max(), attr_tol, int((d > attr_tol).sum())))
            elif not np.array_equal(x, y):
                bad.append("frame %d %s differs at %d of %d" % (f, name, int((x != y).sum()), x.size))
    return bad


def _product_encode_gof(self, frames, params, occupancy_precision=4, stop_after=0, fetch=None):
    """libpccb200 GOF entry point; returns list of GofFrame like Oracle.encode_gof. fetch: iterable of GOF_* ids (default all)."""
    L = self.lib
    L.pccb200_encode_gof.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_i16p), C.POINTER(c_u8p), C.POINTER(C.c_size_t), C.POINTER(SegParams),
                                     C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.pccb200_gof_free.argtypes = [C.c_void_p]
    L.pccb200_gof_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.pccb200_gof_patches.restype = C.c_void_p
    L.pccb200_gof_patches.argtypes = [C.c_void_p, C.c_int]
    L.pccb200_gof_get.restype = C.c_size_t
    L.pccb200_gof_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    n, xs, cs, xp, cp, ns = _frames_args(frames)
    h = C.c_void_p()
    self._check(L.pccb200_encode_gof(self.ctx, n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after, C.byref(h)))
    out = []
    for f in range(n):
        fr = GofFrame()
        w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
        L.pccb200_gof_dims(h, f, C.byref(w), C.byref(hh), C.byref(r))
        fr.width, fr.height = w.value, hh.value
        pl = L.pccb200_gof_patches(h, f)
        cnt = L.pccb200_patches_count(pl)
        patches = np.zeros(cnt, dtype=PATCH_DTYPE)
        depth = np.zeros(L.pccb200_patches_depth_elems(pl), np.int16)
        occ = np.zeros(L.pccb200_patches_occ_elems(pl), np.uint8)
        self._check(L.pccb200_patches_get(pl, patches.ctypes.data_as(C.c_void_p), ptr(depth, c_i16p), ptr(occ, c_u8p)))
        fr.patches = PatchSet(patches, depth, occ)
        for what, dt in GOF_DTYPES.items():
            if fetch is not None and what not in fetch:
                continue
            cnt = L.pccb200_gof_get(h, f, what, None)
            a = np.zeros(cnt, dt)
            if cnt:
                L.pccb200_gof_get(h, f, what, a.ctypes.data_as(C.c_void_p))
            fr.data[what] = a
        out.append(fr)
    L.pccb200_gof_free(h)
    return out


Product.encode_gof = _product_encode_gof


class ProductGof:
    """staged use of the GOF entry points (pack -> [all-reduce canvas] -> resume -> fetch hand-off products)"""

    def __init__(self, product, frames, params, occupancy_precision=4, stop_after=1):
        self.p, L = product, product.lib
        _product_encode_gof.__doc__  # (signatures are declared there)
        L.pccb200_encode_gof.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_i16p), C.POINTER(c_u8p), C.POINTER(C.c_size_t), C.POINTER(SegParams),
                                         C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pccb200_gof_free.argtypes = [C.c_void_p]
        L.pccb200_gof_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.pccb200_gof_get.restype = C.c_size_t
        L.pccb200_gof_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pccb200_gof_resume.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        self.n, self._xs, self._cs, xp, cp, ns = _frames_args(frames)
        self.h = C.c_void_p()
        product._check(L.pccb200_encode_gof(product.ctx, self.n, xp, cp, ns, C.byref(params), occupancy_precision, stop_after, C.byref(self.h)))

    def dims(self, f=0):
        w, hh, r = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self.p.lib.pccb200_gof_dims(self.h, f, C.byref(w), C.byref(hh), C.byref(r))
        return w.value, hh.value, r.value

    def resume(self, width, height, stop_after=0):
        self.p._check(self.p.lib.pccb200_gof_resume(self.h, width, height, stop_after))

    def count(self, f, what):
        return self.p.lib.pccb200_gof_get(self.h, f, what, None)

    def fetch(self, f, what, out=None):
        cnt = self.p.lib.pccb200_gof_get(self.h, f, what, None)
        if out is None or out.size != cnt:
            out = np.empty(cnt, GOF_DTYPES.get(what) or GOF_EXTRA_DTYPES[what])
        if cnt:
            self.p.lib.pccb200_gof_get(self.h, f, what, out.ctypes.data_as(C.c_void_p))
        return out

    def free(self):
        if self.h:
            self.p.lib.pccb200_gof_free(self.h)
            self.h = C.c_void_p()

    def patch_records(self, f):
        """(patch records, block occupancies) of local frame f - what a sharded random-access GOF exchanges (no depth maps)"""
        L = self.p.lib
        L.pccb200_gof_patches.restype = C.c_void_p
        L.pccb200_gof_patches.argtypes = [C.c_void_p, C.c_int]
        pl = L.pccb200_gof_patches(self.h, f)
        patches = np.zeros(L.pccb200_patches_count(pl), dtype=PATCH_DTYPE)
        occ = np.zeros(L.pccb200_patches_occ_elems(pl), np.uint8)
        self.p._check(L.pccb200_patches_get(pl, patches.ctypes.data_as(C.c_void_p), None, ptr(occ, c_u8p)))
        return patches, occ

    def patch_list(self, f):
        """PatchSet of local frame f (records, depth maps, occupancies) as pccb200_gof_patches hands it out"""
        L = self.p.lib
        L.pccb200_gof_patches.restype = C.c_void_p
        L.pccb200_gof_patches.argtypes = [C.c_void_p, C.c_int]
        pl = L.pccb200_gof_patches(self.h, f)
        patches = np.zeros(L.pccb200_patches_count(pl), dtype=PATCH_DTYPE)
        depth = np.zeros(L.pccb200_patches_depth_elems(pl), np.int16)
        occ = np.zeros(L.pccb200_patches_occ_elems(pl), np.uint8)
        self.p._check(L.pccb200_patches_get(pl, patches.ctypes.data_as(C.c_void_p), ptr(depth, c_i16p), ptr(occ, c_u8p)))
        return PatchSet(patches, depth, occ)

    def pack_ra(self, records, local_of):
        """records: [(patch records, occupancies)] of ALL frames of the GOF in frame order; local_of[f]: local index of frame f or -1"""
        L = self.p.lib
        L.pccb200_gof_pack_ra.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, c_u8p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        total = len(records)
        counts = (C.c_int * total)(*[len(r[0]) for r in records])
        sizes = (C.c_size_t * total)(*[len(r[1]) for r in records])
        local = (C.c_int * total)(*[int(x) for x in local_of])
        patches = np.concatenate([np.ascontiguousarray(r[0], PATCH_DTYPE) for r in records]) if total else np.zeros(0, PATCH_DTYPE)
        occ = np.concatenate([np.ascontiguousarray(r[1], np.uint8) for r in records] + [np.zeros(1, np.uint8)])
        self.p._check(L.pccb200_gof_pack_ra(self.h, total, counts, patches.ctypes.data_as(C.c_void_p), ptr(occ, c_u8p), sizes, local))


def _product_generate_point_cloud(self, patches, occ_video, geo0, geo1, width, height, occupancy_precision=4):
    """decoder-side PCCCodec::generatePointCloud: returns dict(xyz, point_to_pixel, partition, boundary)"""
    L = self.lib
    L.pccb200_generate_point_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_u8p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.c_size_t,
                                               C.c_size_t, C.c_int, C.c_size_t, c_i16p, c_u32p, c_u32p, C.POINTER(C.c_uint16), C.POINTER(C.c_size_t)]
    patches = np.ascontiguousarray(patches)
    occ_video, geo0, geo1 = (np.ascontiguousarray(a) for a in (occ_video, geo0, geo1))
    u16p = C.POINTER(C.c_uint16)
    r = C.c_size_t(0)
    args = (self.ctx, patches.ctypes.data_as(C.c_void_p), len(patches), ptr(occ_video, c_u8p), ptr(geo0, u16p), ptr(geo1, u16p), width, height,
            occupancy_precision)
    self._check(L.pccb200_generate_point_cloud(*args, 0, None, None, None, None, C.byref(r)))
    R = r.value
    xyz, p2p = np.zeros((R, 3), np.int16), np.zeros((R, 3), np.uint32)
    part, bnd = np.zeros(R, np.uint32), np.zeros(R, np.uint16)
    self._check(L.pccb200_generate_point_cloud(*args, R, ptr(xyz, c_i16p), ptr(p2p, c_u32p), ptr(part, c_u32p), ptr(bnd, u16p), C.byref(r)))
    return dict(xyz=xyz, point_to_pixel=p2p, partition=part, boundary=bnd)


def _gof_set_decoded(self, f, occ_video=None, geo0=None, geo1=None):
    L = self.p.lib
    u16p = C.POINTER(C.c_uint16)
    L.pccb200_gof_set_decoded.argtypes = [C.c_void_p, C.c_int, c_u8p, u16p, u16p]
    keep = [None if a is None else np.ascontiguousarray(a) for a in (occ_video, geo0, geo1)]
    self.p._check(L.pccb200_gof_set_decoded(self.h, f, None if keep[0] is None else ptr(keep[0], c_u8p),
                                            None if keep[1] is None else ptr(keep[1], u16p), None if keep[2] is None else ptr(keep[2], u16p)))


Product.generate_point_cloud = _product_generate_point_cloud
ProductGof.set_decoded = _gof_set_decoded
