"""ctypes binding of oracle/_ref/libapd_ref.so = the reference's own APD.cu compiled as-is (TEST INFRASTRUCTURE;
needs a GPU; built by `make -C oracle ref` where /root/reference is mounted, shipped to the GPU box prebuilt)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libapd_ref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_eval_costs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_run_pass.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_double)]
        L.ref_run_stages.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 14 + [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_set_sa_mask.argtypes = [C.c_void_p, C.c_longlong]
        L.ref_set_sa_mask.restype = None
        assert L.ref_sizeof_camera() == 120
        _lib = L
    return _lib


def _params(p):
    ip = np.array([p.max_iterations, p.num_images, p.top_k, p.geom_consistency, p.use_impetus, p.strong_radius,
                   p.strong_increment, p.weak_radius, p.weak_increment, p.use_APD, p.use_sa, p.weak_peak_radius,
                   p.rotate_time, p.state], np.int32)
    fp = np.array([p.depth_min, p.depth_max, p.ransac_threshold, p.geom_factor], np.float32)
    return ip, fp


def _ptr_array(arrs):
    if arrs is None:
        return None, None
    keep = [np.ascontiguousarray(a, np.float32) for a in arrs]
    pa = (C.c_void_p * len(keep))(*[a.ctypes.data for a in keep])
    return pa, keep


def _cams(cameras):
    buf = (C.c_char * (120 * len(cameras)))()
    for i, cam in enumerate(cameras):
        C.memmove(C.addressof(buf) + 120 * i, C.byref(cam), 120)
    return buf


def _p(a):
    return None if a is None else a.ctypes.data


def set_sa_mask(labels):
    """segment labels (uint8 [h, w] at the working size) for the following eval_costs / run_pass calls; None clears them"""
    if labels is None:
        lib().ref_set_sa_mask(None, 0)
    else:
        a = np.ascontiguousarray(labels, np.uint8)
        lib().ref_set_sa_mask(a.ctypes.data, a.size)


def eval_costs(images, cameras, params, tuples, planes, mode=0, depths=None, weak=None, selected_views=None, anchors=None):
    h, w = images[0].shape
    n = len(images)
    pi, keep_i = _ptr_array(images)
    pd, keep_d = _ptr_array(depths)
    ip, fp = _params(params)
    ip[1] = n
    tuples = np.ascontiguousarray(tuples, np.int32).reshape(-1, 3)
    planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 4)
    out = np.zeros(len(tuples), np.float32)
    weak = None if weak is None else np.ascontiguousarray(weak, np.uint8)
    sel = None if selected_views is None else np.ascontiguousarray(selected_views, np.uint32)
    anc = None if anchors is None else np.ascontiguousarray(anchors, np.int16)
    cams = _cams(cameras)
    rc = lib().ref_eval_costs(w, h, n, C.cast(pi, C.c_void_p), C.cast(pd, C.c_void_p) if pd else None, C.cast(cams, C.c_void_p),
                              _p(ip), _p(fp), _p(weak), _p(sel), _p(anc), len(tuples), _p(tuples), _p(planes), mode, _p(out))
    if rc != 0:
        raise RuntimeError("ref_eval_costs failed: %d" % rc)
    return out


def run_pass(images, cameras, params, planes=None, weak=None, conf=None, depths=None, seed=1):
    """returns (planes [h,w,4] world normal + depth, weak, conf, RunPatchMatch ms)"""
    h, w = images[0].shape
    n = len(images)
    pi, keep_i = _ptr_array(images)
    pd, keep_d = _ptr_array(depths)
    ip, fp = _params(params)
    ip[1] = n
    planes = np.zeros((h, w, 4), np.float32) if planes is None else np.ascontiguousarray(planes, np.float32).copy()
    weak = np.ones((h, w), np.uint8) if weak is None else np.ascontiguousarray(weak, np.uint8).copy()
    conf = np.ones((h, w), np.uint8) if conf is None else np.ascontiguousarray(conf, np.uint8).copy()
    ms = C.c_double()
    cams = _cams(cameras)
    rc = lib().ref_run_pass(w, h, n, C.cast(pi, C.c_void_p), C.cast(pd, C.c_void_p) if pd else None, C.cast(cams, C.c_void_p),
                            _p(ip), _p(fp), _p(planes), _p(weak), _p(conf), seed, C.byref(ms))
    if rc != 0:
        raise RuntimeError("ref_run_pass failed: %d" % rc)
    return planes, weak, conf, ms.value


STAGES = {"nearest_strong": 0, "neighbour_update": 1, "init": 2, "depth_normal": 3, "median_black": 4, "median_red": 5,
          "depth_to_weak": 6, "confidence": 7, "local_refine": 8}


def run_stages(images, cameras, params, stages, state, depths=None, anchors=None, want_curve=False):
    """Launch the reference's own RNG-free kernels (names of STAGES, in order) on `state` = dict with any of planes [h,w,4],
    costs [h,w], selected_views [h,w] u32, view_weight [h,w,32] u8, weak_info / confidence / weak_reliable [h,w] u8,
    nearest_strong [h,w,2] i16.  Returns a dict of the same arrays after the kernels (+ "curve" [h,w,61])."""
    h, w = images[0].shape
    n = len(images)
    pi, keep_i = _ptr_array(images)
    pd, keep_d = _ptr_array(depths)
    ip, fp = _params(params)
    ip[1] = n
    spec = (("planes", np.float32, (h, w, 4)), ("costs", np.float32, (h, w)), ("selected_views", np.uint32, (h, w)),
            ("view_weight", np.uint8, (h, w, 32)), ("weak_info", np.uint8, (h, w)), ("confidence", np.uint8, (h, w)),
            ("weak_reliable", np.uint8, (h, w)), ("nearest_strong", np.int16, (h, w, 2)))
    out = {}
    for name, dt, shape in spec:
        if name in state and state[name] is not None:
            out[name] = np.ascontiguousarray(state[name], dt).reshape(shape).copy()
        elif name == "planes":
            out[name] = np.zeros(shape, dt)
        else:
            out[name] = None
    anc = None if anchors is None else np.ascontiguousarray(anchors, np.int16)
    st = np.array([STAGES[s] for s in stages], np.int32)
    curve = np.zeros((h, w, 61), np.float32) if want_curve else None
    cams = _cams(cameras)
    rc = lib().ref_run_stages(w, h, n, C.cast(pi, C.c_void_p), C.cast(pd, C.c_void_p) if pd else None, C.cast(cams, C.c_void_p),
                              _p(ip), _p(fp), _p(out["planes"]), _p(out["costs"]), _p(out["selected_views"]), _p(out["view_weight"]),
                              _p(out["weak_info"]), _p(out["confidence"]), _p(out["weak_reliable"]), _p(out["nearest_strong"]), _p(anc),
                              len(st), _p(st), _p(curve))
    if rc != 0:
        raise RuntimeError("ref_run_stages failed: %d" % rc)
    if want_curve:
        out["curve"] = curve
    return out
