"""Shared test helpers: build oracle problems from synthetic scenes (CPU) and mirror them on the GPU library."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import binding as orc  # noqa: E402  (tests are allowed to use the oracle)


def ref_params(state=0, geom=0, use_apd=0, **kw):
    """PatchMatchParams defaults of main.h:80-100 as an oracle struct"""
    p = orc.OParams()
    p.max_iterations, p.num_images, p.top_k = 3, 5, 4
    p.depth_min, p.depth_max = 0.0, 1.0
    p.geom_consistency, p.use_impetus = geom, 1
    p.strong_radius, p.strong_increment, p.weak_radius, p.weak_increment = 5, 2, 5, 5
    p.use_APD, p.use_sa, p.weak_peak_radius, p.rotate_time = use_apd, 1, 2, 4
    p.ransac_threshold, p.geom_factor, p.state = 0.005, 0.2, state
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def oracle_problem(scene, ref, params, depths=None, seed=1, tex_mode=1, threads=8):
    """working resolution == scene resolution (scale_size 1); depth range as APD.cpp:554-555"""
    ids = [ref] + list(scene.pairs[ref])
    imgs = [scene.images[i].astype(np.float32) for i in ids]
    cams = [scene.cameras[i] for i in ids]
    p = orc.OParams()
    C.memmove(C.byref(p), C.byref(params), C.sizeof(p))
    p.depth_min = np.float32(cams[0].depth_min) * np.float32(0.6)
    p.depth_max = np.float32(cams[0].depth_max) * np.float32(1.2)
    d = None if depths is None else [depths[i] for i in ids]
    return orc.Problem(imgs, cams, p, depths=d, seed=seed, stream=ref, tex_mode=tex_mode, num_threads=threads)


def to_apde_params(op):
    from apde_mvs_b200.binding import Params
    p = Params()
    C.memmove(C.byref(p), C.byref(op), C.sizeof(p))
    return p


def oracle_from_ctx(ctx, seed, ref_view, tex_mode=1, threads=8):
    """oracle problem on EXACTLY the inputs of the active GPU problem (working-resolution images, cameras, depth maps,
    parameters), so that every difference is the kernels'."""
    from apde_mvs_b200.binding import FIELD
    w, h, n = ctx.problem_dims()
    cams, prm = ctx.problem_cameras()
    imgs = [ctx.problem_image(i) for i in range(n)]
    op = orc.OParams()
    C.memmove(C.byref(op), C.byref(prm), C.sizeof(op))
    depths = None
    if prm.geom_consistency or prm.use_APD:
        d = ctx.problem_get(FIELD.SRC_DEPTH)
        depths = [d[i] for i in range(n)]
    return orc.Problem(imgs, [cams[i] for i in range(n)], op, depths=depths, seed=seed, stream=ref_view,
                       tex_mode=tex_mode, num_threads=threads)


STATE_FIELDS = ("planes", "costs", "selected_views", "view_weight", "weak_info", "confidence", "fit_planes",
                "weak_reliable", "nearest_strong", "anchors")


def push_state(ctx, pb, fields=STATE_FIELDS):
    """oracle state -> GPU problem"""
    from apde_mvs_b200.binding import FIELD
    m = {"planes": FIELD.PLANES, "costs": FIELD.COSTS, "selected_views": FIELD.SELECTED_VIEWS,
         "view_weight": FIELD.VIEW_WEIGHT, "weak_info": FIELD.WEAK_INFO, "confidence": FIELD.CONFIDENCE,
         "fit_planes": FIELD.FIT_PLANES, "weak_reliable": FIELD.WEAK_RELIABLE, "nearest_strong": FIELD.NEAREST_STRONG,
         "anchors": FIELD.ANCHORS}
    for f in fields:
        ctx.problem_set(m[f], getattr(pb, f))


def pull_state(ctx):
    from apde_mvs_b200.binding import FIELD
    return {"planes": ctx.problem_get(FIELD.PLANES), "costs": ctx.problem_get(FIELD.COSTS),
            "selected_views": ctx.problem_get(FIELD.SELECTED_VIEWS), "view_weight": ctx.problem_get(FIELD.VIEW_WEIGHT),
            "weak_info": ctx.problem_get(FIELD.WEAK_INFO), "confidence": ctx.problem_get(FIELD.CONFIDENCE),
            "fit_planes": ctx.problem_get(FIELD.FIT_PLANES), "weak_reliable": ctx.problem_get(FIELD.WEAK_RELIABLE),
            "nearest_strong": ctx.problem_get(FIELD.NEAREST_STRONG), "anchors": ctx.problem_get(FIELD.ANCHORS)}


def plane_depth(planes, cam, xs=None, ys=None):
    """ComputeDepthfromPlaneHypothesis (APD.cu:237) vectorised, planes [h, w, 4] in the camera frame"""
    h, w = planes.shape[:2]
    gx, gy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    with np.errstate(all="ignore"):
        return -planes[..., 3] * fx / ((gx - cx) * planes[..., 0] + (fx / fy) * (gy - cy) * planes[..., 1] + fx * planes[..., 2])
