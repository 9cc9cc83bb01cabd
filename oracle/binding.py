"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs only.  The product
(apde_mvs_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libapd_oracle.so")
MAX_IMAGES, ANCHOR_NUM = 32, 9


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("apd_oracle.cpp", "apd_oracle.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libapd_oracle.so"])
    return LIB


class OCamera(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("R", C.c_float * 9), ("t", C.c_float * 3), ("c", C.c_float * 3),
                ("height", C.c_int), ("width", C.c_int), ("depth_min", C.c_float), ("depth_max", C.c_float),
                ("interval", C.c_float), ("depth_num", C.c_float)]


class OParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("num_images", C.c_int), ("top_k", C.c_int), ("depth_min", C.c_float),
                ("depth_max", C.c_float), ("geom_consistency", C.c_int), ("use_impetus", C.c_int),
                ("strong_radius", C.c_int), ("strong_increment", C.c_int), ("weak_radius", C.c_int),
                ("weak_increment", C.c_int), ("use_APD", C.c_int), ("use_sa", C.c_int), ("weak_peak_radius", C.c_int),
                ("rotate_time", C.c_int), ("ransac_threshold", C.c_float), ("geom_factor", C.c_float), ("state", C.c_int)]


class OProblem(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("num_images", C.c_int),
                ("images", C.c_void_p * MAX_IMAGES), ("depths", C.c_void_p * MAX_IMAGES),
                ("cameras", OCamera * MAX_IMAGES), ("params", OParams),
                ("planes", C.c_void_p), ("costs", C.c_void_p), ("selected_views", C.c_void_p),
                ("view_weight", C.c_void_p), ("weak_info", C.c_void_p), ("confidence", C.c_void_p),
                ("fit_planes", C.c_void_p), ("weak_reliable", C.c_void_p), ("nearest_strong", C.c_void_p),
                ("anchors", C.c_void_p), ("seed", C.c_uint32), ("stream", C.c_uint32), ("tex_mode", C.c_int),
                ("num_threads", C.c_int), ("counters", C.c_uint64 * 8), ("sa_mask", C.c_void_p)]


class OFusionInput(C.Structure):
    _fields_ = [("num_views", C.c_int), ("width", C.c_int), ("height", C.c_int), ("cameras", C.c_void_p),
                ("depths", C.c_void_p), ("normals", C.c_void_p), ("weaks", C.c_void_p), ("confidences", C.c_void_p),
                ("colors", C.c_void_p), ("src_offsets", C.c_void_p), ("src_ids", C.c_void_p), ("num_threads", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        PP = C.POINTER(OProblem)
        L.orc_ncc_old.restype = L.orc_ncc_new.restype = L.orc_geom_cost.restype = C.c_float
        for f in (L.orc_ncc_old, L.orc_ncc_new, L.orc_geom_cost):
            f.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_tex2d.restype = C.c_float
        L.orc_tex2d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_eval_costs.argtypes = [PP, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_homography.argtypes = [C.POINTER(OCamera), C.POINTER(OCamera), C.c_void_p, C.c_void_p]
        for name in ("orc_nearest_strong", "orc_gen_anchors", "orc_neighbour_update", "orc_random_init",
                     "orc_depth_normal", "orc_confidence", "orc_local_refine", "orc_run_pass"):
            getattr(L, name).argtypes = [PP]
        L.orc_propagate_strong.argtypes = [PP, C.c_int, C.c_int]
        L.orc_propagate_weak.argtypes = [PP, C.c_int, C.c_int]
        L.orc_ransac_fit.argtypes = [PP, C.c_int]
        L.orc_median_filter.argtypes = [PP, C.c_int]
        L.orc_depth_to_weak.argtypes = [PP, C.c_void_p]
        L.orc_checkerboard_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_philox.argtypes = [C.c_uint32] * 5 + [C.c_void_p]
        L.orc_weak_vis_filter.argtypes = [C.POINTER(OFusionInput), C.c_void_p]
        L.orc_fuse.restype = C.c_int64
        L.orc_fuse.argtypes = [C.POINTER(OFusionInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_fuse_tat.restype = C.c_int64
        L.orc_fuse_tat.argtypes = [C.POINTER(OFusionInput), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def _copy_struct(dst, src):
    C.memmove(C.byref(dst), C.byref(src), C.sizeof(dst))


class Problem:
    """One reference view + its sources at working resolution, with all per-pixel state as numpy arrays."""

    def __init__(self, images, cameras, params, depths=None, seed=1, stream=0, tex_mode=1, num_threads=8):
        self.L = lib()
        self.images = [np.ascontiguousarray(im, np.float32) for im in images]
        self.h, self.w = self.images[0].shape
        n = len(self.images)
        P = self.h * self.w
        self.pb = OProblem()
        pb = self.pb
        pb.width, pb.height, pb.num_images = self.w, self.h, n
        for i in range(n):
            pb.images[i] = self.images[i].ctypes.data
            _copy_struct(pb.cameras[i], cameras[i])
        self.depths = None
        if depths is not None:
            self.depths = [np.ascontiguousarray(d, np.float32) for d in depths]
            for i in range(n):
                pb.depths[i] = self.depths[i].ctypes.data
        _copy_struct(pb.params, params)
        pb.params.num_images = n
        self.planes = np.zeros((self.h, self.w, 4), np.float32)
        self.costs = np.zeros((self.h, self.w), np.float32)
        self.selected_views = np.zeros((self.h, self.w), np.uint32)
        self.view_weight = np.zeros((self.h, self.w, 32), np.uint8)
        self.weak_info = np.full((self.h, self.w), 1, np.uint8)
        self.confidence = np.ones((self.h, self.w), np.uint8)
        self.fit_planes = np.zeros((self.h, self.w, 4), np.float32)
        self.weak_reliable = np.zeros((self.h, self.w), np.uint8)
        self.nearest_strong = np.full((self.h, self.w, 2), -1, np.int16)
        self.anchors = np.full((self.h, self.w, ANCHOR_NUM * 2), -1, np.int16)
        for name in ("planes", "costs", "selected_views", "view_weight", "weak_info", "confidence", "fit_planes",
                     "weak_reliable", "nearest_strong", "anchors"):
            setattr(pb, name, getattr(self, name).ctypes.data)
        pb.seed, pb.stream, pb.tex_mode, pb.num_threads = seed, stream, tex_mode, num_threads
        assert P > 0

    def set_sa_mask(self, labels):
        """segment labels of the reference view at the working size (uint8 [h, w]); None = all zero (APD.cpp:613, 641-649)"""
        self.sa_mask = None if labels is None else np.ascontiguousarray(labels, np.uint8)
        if self.sa_mask is not None:
            assert self.sa_mask.shape == (self.h, self.w)
        self.pb.sa_mask = None if self.sa_mask is None else self.sa_mask.ctypes.data

    @property
    def ptr(self):
        return C.byref(self.pb)

    def eval_costs(self, tuples, planes, mode=0):
        tuples = np.ascontiguousarray(tuples, np.int32).reshape(-1, 3)
        planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 4)
        out = np.zeros(len(tuples), np.float32)
        self.L.orc_eval_costs(self.ptr, len(tuples), tuples.ctypes.data, planes.ctypes.data, mode, out.ctypes.data)
        return out

    def stage(self, name, *args):
        getattr(self.L, "orc_" + name)(self.ptr, *args)

    def counters(self):
        return [int(x) for x in self.pb.counters[:3]]


def philox(seed, stream, pixel, site, block):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(seed, stream, pixel, site, block, out.ctypes.data)
    return out


def checkerboard_candidates(costs, x, y):
    costs = np.ascontiguousarray(costs, np.float32)
    h, w = costs.shape
    pos = np.zeros(8, np.int32)
    flags = np.zeros(8, np.uint8)
    lib().orc_checkerboard_candidates(costs.ctypes.data, w, h, x, y, pos.ctypes.data, flags.ctypes.data)
    return pos, flags


def fusion(cameras, depths, normals, weaks, confs, pairs, colors=None, weak_filter=True, num_threads=8, variant=0):
    """returns (points xyz, colours bgr, skip_weaks); variant 0 = RunFusion, 1 = RunFusion_TAT_I, 2 = RunFusion_TAT_A"""
    L = lib()
    V, h, w = depths.shape
    cams = (OCamera * V)()
    for i in range(V):
        _copy_struct(cams[i], cameras[i])
    depths = np.ascontiguousarray(depths, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    weaks = np.ascontiguousarray(weaks, np.uint8)
    confs = np.ascontiguousarray(confs, np.uint8)
    offs = np.zeros(V + 1, np.int32)
    ids = []
    for i, p in enumerate(pairs):
        ids += list(p)
        offs[i + 1] = len(ids)
    ids = np.ascontiguousarray(ids, np.int32)
    fi = OFusionInput()
    fi.num_views, fi.width, fi.height = V, w, h
    fi.cameras = C.cast(cams, C.c_void_p)
    fi.depths, fi.normals, fi.weaks, fi.confidences = depths.ctypes.data, normals.ctypes.data, weaks.ctypes.data, confs.ctypes.data
    if colors is not None:
        colors = np.ascontiguousarray(colors, np.uint8)
        fi.colors = colors.ctypes.data
    fi.src_offsets, fi.src_ids, fi.num_threads = offs.ctypes.data, ids.ctypes.data, num_threads
    skip = np.zeros((V, h, w), np.uint8)
    if weak_filter:
        L.orc_weak_vis_filter(C.byref(fi), skip.ctypes.data)
    def run(p, c, m):
        if variant == 0:
            return L.orc_fuse(C.byref(fi), skip.ctypes.data, p, c, m)
        return L.orc_fuse_tat(C.byref(fi), skip.ctypes.data, variant, p, c, m)
    n = run(None, None, 0)
    xyz = np.zeros((n, 3), np.float32)
    bgr = np.zeros((n, 3), np.float32)
    run(xyz.ctypes.data, bgr.ctypes.data, n)
    return xyz, bgr, skip
