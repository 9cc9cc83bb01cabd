"""ctypes mirror of include/apde.h (host buffers in, host buffers out; numpy only for the buffers)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_IMAGES = 32
ANCHOR_NUM = 9
COMM_ID_BYTES = 128


def lib_path():
    # APDE_LIB: an alternative build of the same ABI (A/B comparisons of library versions, tools/ab_prev.sh)
    return os.environ.get("APDE_LIB") or os.path.join(_HERE, "_build", "libapde.so")


class ApdeError(RuntimeError):
    pass


class Camera(C.Structure):
    """apde_camera == reference Camera (main.h:50-61)."""

    _fields_ = [
        ("K", C.c_float * 9), ("R", C.c_float * 9), ("t", C.c_float * 3), ("c", C.c_float * 3),
        ("height", C.c_int), ("width", C.c_int),
        ("depth_min", C.c_float), ("depth_max", C.c_float), ("interval", C.c_float), ("depth_num", C.c_float),
    ]

    @staticmethod
    def make(K, R, t, width, height, depth_min, depth_max, depth_num=192.0):
        cam = Camera()
        K = np.asarray(K, np.float32).reshape(9)
        R = np.asarray(R, np.float32).reshape(9)
        t = np.asarray(t, np.float32).reshape(3)
        for i in range(9):
            cam.K[i] = float(K[i])
            cam.R[i] = float(R[i])
        for i in range(3):
            cam.t[i] = float(t[i])
        # camera centre in double, rounded to float (ReadCamera, APD.cpp:114-119)
        R64, t64 = R.astype(np.float64).reshape(3, 3), t.astype(np.float64)
        c = -(R64.T @ t64)
        for i in range(3):
            cam.c[i] = float(np.float32(c[i]))
        cam.width, cam.height = int(width), int(height)
        cam.depth_min, cam.depth_max = float(depth_min), float(depth_max)
        cam.depth_num = float(depth_num)
        cam.interval = float((depth_max - depth_min) / (depth_num - 1.0))
        return cam


class Params(C.Structure):
    """apde_params == reference PatchMatchParams (main.h:80-100), bools as int."""

    _fields_ = [
        ("max_iterations", C.c_int), ("num_images", C.c_int), ("top_k", C.c_int),
        ("depth_min", C.c_float), ("depth_max", C.c_float),
        ("geom_consistency", C.c_int), ("use_impetus", C.c_int),
        ("strong_radius", C.c_int), ("strong_increment", C.c_int), ("weak_radius", C.c_int), ("weak_increment", C.c_int),
        ("use_APD", C.c_int), ("use_sa", C.c_int), ("weak_peak_radius", C.c_int), ("rotate_time", C.c_int),
        ("ransac_threshold", C.c_float), ("geom_factor", C.c_float), ("state", C.c_int),
    ]


class Schedule(C.Structure):
    _fields_ = [
        ("rounds", C.c_int), ("geom_iterations", C.c_int), ("jacobi", C.c_int), ("use_impetus", C.c_int),
        ("geom_factor", C.c_float), ("seed", C.c_uint32), ("first_view", C.c_int), ("num_views_local", C.c_int),
        ("use_sa", C.c_int),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("patchmatch_ms", C.c_double), ("device_ms", C.c_double), ("total_ms", C.c_double),
        ("evals_ncc_old", C.c_uint64), ("evals_ncc_new", C.c_uint64), ("evals_geom", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("passes", C.c_int), ("pad_", C.c_int),
        ("exchange_ms", C.c_double), ("exchange_bytes", C.c_uint64),
    ]


class STAGE:
    NEAREST_STRONG, GEN_ANCHORS, INIT, PROP_STRONG, RANSAC_FIT, PROP_WEAK, DEPTH_NORMAL, MEDIAN, DEPTH_TO_WEAK, \
        CONFIDENCE, LOCAL_REFINE = range(11)


class POOL:
    DEPTH, NORMAL, WEAK, CONFIDENCE, SKIP = range(5)


class FIELD:
    PLANES, COSTS, SELECTED_VIEWS, VIEW_WEIGHT, WEAK_INFO, CONFIDENCE, FIT_PLANES, WEAK_RELIABLE, NEAREST_STRONG, \
        ANCHORS, IMAGE, SRC_DEPTH = range(12)
    RELIABLE_CURVE, SA_MASK = 12, 13


FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2

_lib = None


def load_library(path=None):
    """dlopen libapde.so; raises if it was not built (there is no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or lib_path()
    if not os.path.exists(path):
        raise ApdeError("libapde.so not built (%s): run `python apde_mvs_b200/build.py`" % path)
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.apde_last_error.restype = C.c_char_p
    lib.apde_version.restype = C.c_char_p
    lib.apde_params_default.argtypes = [C.POINTER(Params)]
    lib.apde_schedule_default.argtypes = [C.POINTER(Schedule)]
    lib.apde_create.argtypes = [C.c_int, C.POINTER(P)]
    lib.apde_destroy.argtypes = [P]
    lib.apde_destroy.restype = None
    lib.apde_scene_begin.argtypes = [P, C.c_int, C.c_int, C.c_int]
    lib.apde_scene_set_view.argtypes = [P, C.c_int, P, P, C.POINTER(Camera)]
    lib.apde_scene_set_pairs.argtypes = [P, C.c_int, C.c_int, P]
    lib.apde_view_set_sa_mask.argtypes = [P, C.c_int, P, C.c_int, C.c_int]
    lib.apde_scene_commit.argtypes = [P]
    lib.apde_view_download.argtypes = [P, C.c_int, P, P, P, P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.apde_view_upload.argtypes = [P, C.c_int, P, P, P, P, C.c_int, C.c_int]
    lib.apde_problem_setup.argtypes = [P, C.c_int, C.POINTER(Params), C.c_int, C.c_uint32]
    lib.apde_problem_stage.argtypes = [P, C.c_int, C.c_int, C.c_int]
    lib.apde_problem_run.argtypes = [P]
    lib.apde_problem_get.argtypes = [P, C.c_int, P, C.c_size_t]
    lib.apde_problem_set.argtypes = [P, C.c_int, P, C.c_size_t]
    lib.apde_problem_dims.argtypes = [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.apde_problem_get_image.argtypes = [P, C.c_int, P, C.c_size_t]
    lib.apde_problem_get_cameras.argtypes = [P, C.POINTER(Camera), C.POINTER(Params)]
    lib.apde_problem_finish.argtypes = [P]
    lib.apde_pass_run.argtypes = [P, C.c_int, C.POINTER(Params), C.c_int, C.c_uint32]
    lib.apde_eval_costs.argtypes = [P, C.c_int, P, P, C.c_int, P]
    lib.apde_debug_tex2d.argtypes = [P, C.c_int, C.c_int, P, P]
    lib.apde_run_schedule.argtypes = [P, C.POINTER(Schedule), C.POINTER(Timing)]
    lib.apde_run_schedule_pass.argtypes = [P, C.POINTER(Schedule), C.c_int, C.POINTER(Timing)]
    lib.apde_schedule_num_passes.argtypes = [P, C.POINTER(Schedule)]
    lib.apde_schedule_pass_params.argtypes = [P, C.POINTER(Schedule), C.c_int, C.POINTER(Params), C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    lib.apde_problem_capture_curve.argtypes = [P, C.c_int]
    lib.apde_fuse_take_points.argtypes = [P, P, P, C.c_int64, C.POINTER(C.c_int64)]
    lib.apde_get_anchor_evals.argtypes = [P, C.POINTER(C.c_uint64)]
    lib.apde_get_counters.argtypes = [P, C.POINTER(C.c_uint64), C.c_int]
    lib.apde_set_profiling.argtypes = [P, C.c_int]
    lib.apde_set_sweep_budget_mb.argtypes = [P, C.c_size_t]
    lib.apde_get_stage_stats.argtypes = [P, P, P, P, C.c_int]
    lib.apde_microbench.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.apde_depth_pool.argtypes = [P, C.POINTER(P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.apde_weak_vis_filter.argtypes = [P, P]
    lib.apde_fuse.argtypes = [P, C.c_int, P, P, C.c_int64, C.POINTER(C.c_int64)]
    lib.apde_fuse_variant.argtypes = [P, C.c_int, C.c_int, P, P, C.c_int64, C.POINTER(C.c_int64)]
    lib.apde_weak_vis_filter_range.argtypes = [P, C.c_int, C.c_int, P]
    lib.apde_map_pool.argtypes = [P, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.apde_views_mark_maps.argtypes = [P, C.c_int, C.c_int]
    lib.apde_comm_create_id.argtypes = [P]
    lib.apde_comm_init.argtypes = [P, P, C.c_int, C.c_int]
    lib.apde_comm_block_of.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.apde_comm_info.argtypes = [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.apde_exchange.argtypes = [P, C.c_int]
    lib.apde_fuse_collective.argtypes = [P, C.c_int, C.c_int, P, P, C.c_int64, C.POINTER(C.c_int64)]
    lib.apde_comm_destroy.argtypes = [P]
    _lib = lib
    return lib


def default_params():
    p = Params()
    load_library().apde_params_default(C.byref(p))
    return p


def default_schedule():
    s = Schedule()
    load_library().apde_schedule_default(C.byref(s))
    return s


_FIELD_DTYPE = {
    FIELD.PLANES: (np.float32, 4), FIELD.COSTS: (np.float32, 1), FIELD.SELECTED_VIEWS: (np.uint32, 1),
    FIELD.VIEW_WEIGHT: (np.uint8, 32), FIELD.WEAK_INFO: (np.uint8, 1), FIELD.CONFIDENCE: (np.uint8, 1),
    FIELD.FIT_PLANES: (np.float32, 4), FIELD.WEAK_RELIABLE: (np.uint8, 1), FIELD.NEAREST_STRONG: (np.int16, 2),
    FIELD.ANCHORS: (np.int16, 2 * ANCHOR_NUM), FIELD.IMAGE: (np.float32, 1), FIELD.SA_MASK: (np.uint8, 1),
    FIELD.RELIABLE_CURVE: (np.float32, 61),
}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU context (apde_context).  Mirrors the reference driver objects: scene = dense folder contents,
    problem = the APD class (APD.h:88-114), schedule = main()'s loops (main.cpp:303-367)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        self._h = None
        self._check(self.lib.apde_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise ApdeError("libapde error %d: %s" % (rc, self.lib.apde_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.apde_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene
    def scene_begin(self, num_views, width, height):
        self._check(self.lib.apde_scene_begin(self._h, num_views, width, height))
        self.V, self.W, self.H = num_views, width, height

    def scene_set_view(self, view, gray, cam, bgr=None):
        gray = np.ascontiguousarray(gray, np.uint8)
        assert gray.shape == (self.H, self.W)
        if bgr is not None:
            bgr = np.ascontiguousarray(bgr, np.uint8)
            assert bgr.shape == (self.H, self.W, 3)
        self._check(self.lib.apde_scene_set_view(self._h, view, _ptr(gray), _ptr(bgr), C.byref(cam)))

    def scene_set_pairs(self, view, src_views):
        a = np.ascontiguousarray(src_views, np.int32)
        self._check(self.lib.apde_scene_set_pairs(self._h, view, len(a), _ptr(a)))

    def scene_commit(self):
        self._check(self.lib.apde_scene_commit(self._h))

    def load_scene(self, scene):
        """scene: apde_mvs_b200.scene.Scene"""
        self.scene_begin(len(scene.images), scene.width, scene.height)
        for v in range(len(scene.images)):
            self.scene_set_view(v, scene.images[v], scene.cameras[v], scene.colors[v] if scene.colors else None)
            self.scene_set_pairs(v, scene.pairs[v])
        self.scene_commit()

    def view_set_sa_mask(self, view, labels):
        """segment-label map of a view (uint8 [h, w] of any size, as in sa_masks/<id>.bin); None removes it"""
        if labels is None:
            self._check(self.lib.apde_view_set_sa_mask(self._h, view, None, 0, 0))
            return
        a = np.ascontiguousarray(labels, np.uint8)
        self._check(self.lib.apde_view_set_sa_mask(self._h, view, _ptr(a), a.shape[1], a.shape[0]))

    def view_download(self, view):
        w, h = C.c_int(), C.c_int()
        self._check(self.lib.apde_view_download(self._h, view, None, None, None, None, C.byref(w), C.byref(h)))
        P = w.value * h.value
        depth = np.zeros((h.value, w.value), np.float32)
        normal = np.zeros((h.value, w.value, 3), np.float32)
        weak = np.zeros((h.value, w.value), np.uint8)
        conf = np.zeros((h.value, w.value), np.uint8)
        if P:
            self._check(self.lib.apde_view_download(self._h, view, _ptr(depth), _ptr(normal), _ptr(weak), _ptr(conf),
                                                    C.byref(w), C.byref(h)))
        return depth, normal, weak, conf

    def view_dims(self, view):
        """(width, height) of the maps stored for a view, (0, 0) if none yet"""
        w, h = C.c_int(), C.c_int()
        self._check(self.lib.apde_view_download(self._h, view, None, None, None, None, C.byref(w), C.byref(h)))
        return w.value, h.value

    def view_upload(self, view, depth=None, normal=None, weak=None, conf=None):
        ref = next(a for a in (depth, normal, weak, conf) if a is not None)
        h, w = ref.shape[:2]
        depth = None if depth is None else np.ascontiguousarray(depth, np.float32)
        normal = None if normal is None else np.ascontiguousarray(normal, np.float32)
        weak = None if weak is None else np.ascontiguousarray(weak, np.uint8)
        conf = None if conf is None else np.ascontiguousarray(conf, np.uint8)
        self._check(self.lib.apde_view_upload(self._h, view, _ptr(depth), _ptr(normal), _ptr(weak), _ptr(conf), w, h))

    # ---- problem
    def problem_setup(self, ref_view, params, scale_size=1, seed=1):
        self._check(self.lib.apde_problem_setup(self._h, ref_view, C.byref(params), scale_size, seed))

    def problem_dims(self):
        w, h, n = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.apde_problem_dims(self._h, C.byref(w), C.byref(h), C.byref(n)))
        return w.value, h.value, n.value

    def problem_stage(self, stage, iter=0, color=0):
        self._check(self.lib.apde_problem_stage(self._h, stage, iter, color))

    def problem_run(self):
        self._check(self.lib.apde_problem_run(self._h))

    def problem_finish(self):
        self._check(self.lib.apde_problem_finish(self._h))

    def problem_get(self, field):
        w, h, n = self.problem_dims()
        if field == FIELD.SRC_DEPTH:
            a = np.zeros((n, h, w), np.float32)
        else:
            dt, k = _FIELD_DTYPE[field]
            a = np.zeros((h, w, k) if k > 1 else (h, w), dt)
        self._check(self.lib.apde_problem_get(self._h, field, _ptr(a), a.nbytes))
        return a

    def problem_set(self, field, a):
        a = np.ascontiguousarray(a)
        self._check(self.lib.apde_problem_set(self._h, field, _ptr(a), a.nbytes))

    def problem_image(self, idx):
        w, h, _ = self.problem_dims()
        a = np.zeros((h, w), np.float32)
        self._check(self.lib.apde_problem_get_image(self._h, idx, _ptr(a), a.nbytes))
        return a

    def problem_cameras(self):
        _, _, n = self.problem_dims()
        cams = (Camera * n)()
        p = Params()
        self._check(self.lib.apde_problem_get_cameras(self._h, cams, C.byref(p)))
        return cams, p

    def pass_run(self, ref_view, params, scale_size=1, seed=1):
        self._check(self.lib.apde_pass_run(self._h, ref_view, C.byref(params), scale_size, seed))

    def eval_costs(self, tuples, planes, mode=0):
        tuples = np.ascontiguousarray(tuples, np.int32).reshape(-1, 3)
        planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 4)
        assert len(tuples) == len(planes)
        out = np.zeros(len(tuples), np.float32)
        self._check(self.lib.apde_eval_costs(self._h, len(tuples), _ptr(tuples), _ptr(planes), mode, _ptr(out)))
        return out

    def debug_tex2d(self, idx, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros(len(xy), np.float32)
        self._check(self.lib.apde_debug_tex2d(self._h, idx, len(xy), _ptr(xy), _ptr(out)))
        return out

    # ---- schedule
    def schedule_pass_params(self, sched, pass_index):
        """(Params, scale_size, seed) that pass `pass_index` of the schedule uses for every view (main.cpp:309-365)"""
        p, scale, seed = Params(), C.c_int(), C.c_uint32()
        self._check(self.lib.apde_schedule_pass_params(self._h, C.byref(sched), pass_index, C.byref(p), C.byref(scale), C.byref(seed)))
        return p, scale.value, seed.value

    def capture_curve(self, on=True):
        """keep DepthToWeak's 61-sample cost curve of the following problems (FIELD.RELIABLE_CURVE)"""
        self._check(self.lib.apde_problem_capture_curve(self._h, 1 if on else 0))

    def num_passes(self, sched):
        n = self.lib.apde_schedule_num_passes(self._h, C.byref(sched))
        if n < 0:
            self._check(n)
        return n

    def run_schedule(self, sched):
        t = Timing()
        self._check(self.lib.apde_run_schedule(self._h, C.byref(sched), C.byref(t)))
        return t

    def run_schedule_pass(self, sched, pass_index, timing=None):
        t = timing if timing is not None else Timing()
        self._check(self.lib.apde_run_schedule_pass(self._h, C.byref(sched), pass_index, C.byref(t)))
        return t

    def counters(self, reset=False):
        out = (C.c_uint64 * 4)()
        self._check(self.lib.apde_get_counters(self._h, out, 1 if reset else 0))
        return [int(x) for x in out]

    def anchor_evals(self):
        """3x3 anchor patches sampled by the weak propagation since the last counter reset (9 samples each)"""
        out = C.c_uint64(0)
        self._check(self.lib.apde_get_anchor_evals(self._h, C.byref(out)))
        return int(out.value)

    def set_sweep_budget_mb(self, mb):
        self._check(self.lib.apde_set_sweep_budget_mb(self._h, C.c_size_t(mb)))

    def set_profiling(self, on=True):
        self._check(self.lib.apde_set_profiling(self._h, int(on)))

    def stage_stats(self, reset=False):
        ms = np.zeros(11, np.float64)
        launches = np.zeros(11, np.uint64)
        evals = np.zeros((11, 3), np.uint64)
        self._check(self.lib.apde_get_stage_stats(self._h, _ptr(ms), _ptr(launches), _ptr(evals), int(reset)))
        return ms, launches, evals

    def microbench(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.apde_microbench(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def depth_pool(self):
        p, nbytes, per_view = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self.lib.apde_depth_pool(self._h, C.byref(p), C.byref(nbytes), C.byref(per_view)))
        return p.value, nbytes.value, per_view.value

    # ---- fusion
    def weak_vis_filter(self):
        d, _, _, _ = self.view_download(0)
        skip = np.zeros((self.V,) + d.shape, np.uint8)
        self._check(self.lib.apde_weak_vis_filter(self._h, _ptr(skip)))
        return skip

    def weak_vis_filter_range(self, first_view, num_views):
        """WeakVisFilter for a shard of reference views; the skip maps stay in the device pool (POOL.SKIP)"""
        self._check(self.lib.apde_weak_vis_filter_range(self._h, first_view, num_views, None))

    def map_pool(self, which):
        """(device pointer, total bytes, bytes per view) of the [V][...] pool of one map field"""
        ptr, total, per = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self.lib.apde_map_pool(self._h, which, C.byref(ptr), C.byref(total), C.byref(per)))
        return ptr.value, total.value, per.value

    def views_mark_maps(self, width, height):
        self._check(self.lib.apde_views_mark_maps(self._h, width, height))

    # ---- multi-GPU jobs (apde_comm_*): NCCL inside the library, the caller only hands the 128-byte id around
    @staticmethod
    def comm_create_id():
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        lib = load_library()
        if lib.apde_comm_create_id(C.cast(buf, C.c_void_p)) != 0:
            raise ApdeError("apde_comm_create_id: %s" % lib.apde_last_error().decode())
        return bytes(buf)

    def comm_init(self, comm_id, rank, world):
        assert len(comm_id) == COMM_ID_BYTES
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._check(self.lib.apde_comm_init(self._h, C.cast(buf, C.c_void_p), rank, world))

    def comm_info(self):
        """(rank, world, first_view, num_views) of this context in its job (a lone context is rank 0 of 1)"""
        r, w, f, n = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.apde_comm_info(self._h, C.byref(r), C.byref(w), C.byref(f), C.byref(n)))
        return r.value, w.value, f.value, n.value

    def exchange(self, which):
        self._check(self.lib.apde_exchange(self._h, which))

    def fuse_collective(self, use_weak_filter=True, variant=0):
        """collective RunFusion over the maps of all ranks: (xyz, bgr) on rank 0, (None, None) elsewhere"""
        rank = self.comm_info()[0]
        n = C.c_int64()
        flt = 1 if use_weak_filter else 0
        # one collective call; rank 0's context keeps the cloud until take_points copies it out
        self._check(self.lib.apde_fuse_collective(self._h, variant, flt, None, None, 0, C.byref(n)))
        if rank != 0:
            return None, None
        return self._take_points(n.value)

    def _take_points(self, n):
        xyz = np.zeros((n, 3), np.float32)
        bgr = np.zeros((n, 3), np.float32)
        n2 = C.c_int64()
        self._check(self.lib.apde_fuse_take_points(self._h, _ptr(xyz), _ptr(bgr), n, C.byref(n2)))
        assert n2.value == n
        return xyz, bgr

    def fuse(self, use_weak_filter=True, max_points=None, variant=0):
        """variant 0 = RunFusion, 1 = RunFusion_TAT_I, 2 = RunFusion_TAT_A (main.cpp:277-283)"""
        n = C.c_int64()
        if max_points is None:  # fusion runs once: count (the context keeps the cloud), then take
            self._check(self.lib.apde_fuse_variant(self._h, variant, int(use_weak_filter), None, None, 0, C.byref(n)))
            return self._take_points(n.value)
        xyz = np.zeros((max_points, 3), np.float32)
        bgr = np.zeros((max_points, 3), np.float32)
        self._check(self.lib.apde_fuse_variant(self._h, variant, int(use_weak_filter), _ptr(xyz), _ptr(bgr), max_points,
                                               C.byref(n)))
        k = min(n.value, max_points)
        return xyz[:k], bgr[:k]
