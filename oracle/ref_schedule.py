"""The reference's multi-scale schedule (main.cpp:303-367) driven around the reference's OWN kernels (oracle/_ref,
APD::RunPatchMatch compiled unmodified) -- TEST / BASELINE INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).

The host stages either side of RunPatchMatch are emulated with cv2 / numpy exactly as the reference performs them:
image pyramid by cv::resize INTER_LINEAR (APD.cpp:574), map hand-over between rounds by INTER_NEAREST (APD.cpp:607-626,
669-673), camera rescale (APD.cpp:576-588), depth range check (main.cpp:172-175)."""
import ctypes as C

import numpy as np

from . import binding as orc
from . import ref_binding as ref


def default_params():
    """PatchMatchParams defaults of main.h:80-100"""
    p = orc.OParams()
    p.max_iterations, p.num_images, p.top_k = 3, 5, 4
    p.depth_min, p.depth_max = 0.0, 1.0
    p.geom_consistency, p.use_impetus = 0, 1
    p.strong_radius, p.strong_increment, p.weak_radius, p.weak_increment = 5, 2, 5, 5
    p.use_APD, p.use_sa, p.weak_peak_radius, p.rotate_time = 0, 1, 2, 4
    p.ransac_threshold, p.geom_factor, p.state = 0.005, 0.2, 0
    return p


def compute_round_num(width, height):
    """ComputeRoundNum, main.cpp:129-146"""
    rounds, m = 1, max(width, height)
    while m > 800:
        m //= 2
        rounds += 1
    return rounds


def sample_views(V, nviews):
    """`nviews` reference views spread evenly over the scene (the middle of each of nviews equal blocks); all views for 0"""
    if nviews <= 0 or nviews >= V:
        return list(range(V))
    return sorted({min(V - 1, int((i + 0.5) * V / nviews)) for i in range(nviews)})


def run_reference_schedule(scene, rounds=0, geom_iters=3, nviews=0, seed_base=0, camera_type=None, geom_factor=0.2):
    """returns (maps, patchmatch_ms): maps[v] = dict(depth, normal, weak, conf) after the last pass; patchmatch_ms = sum of
    the reference's own 'RunPatchMatch time' over the timed views (`nviews` views spread over the scene, sample_views; 0 = all).
    Every view runs the first photometric pass (the later passes of the timed views read those depth maps)."""
    import cv2
    V = len(scene.images)
    timed = set(sample_views(V, nviews))
    W, H = scene.width, scene.height
    rounds = rounds if rounds > 0 else compute_round_num(W, H)
    fimgs = [im.astype(np.float32) for im in scene.images]
    Camera = camera_type or type(scene.cameras[0])
    maps = [dict() for _ in range(V)]
    pm_ms = 0.0
    it = 0
    for i in range(rounds):
        scale = 2 ** (rounds - 1 - i)
        w, h = int(round(W / scale)), int(round(H / scale))
        lv_imgs = fimgs if scale == 1 else [cv2.resize(im, (w, h), interpolation=cv2.INTER_LINEAR) for im in fimgs]
        for j in range(-1, geom_iters):
            p = default_params()
            p.geom_factor = geom_factor
            p.use_APD = 0 if i == 0 else 1
            if i > 0:
                p.ransac_threshold = 0.01 - i * 0.00125
                p.rotate_time = min(2 ** i, 4)
            if j < 0:
                p.state, p.geom_consistency, p.weak_peak_radius = (0 if i == 0 else 1), 0, 6
            else:
                p.state, p.geom_consistency, p.weak_peak_radius = 2, 1, max(4 - 2 * j, 2)
            for v in range(V):
                if (i, j) != (0, -1) and v not in timed:  # every view needs a depth map after the first pass
                    continue
                ids = [v] + list(scene.pairs[v])
                cams = []
                for k in ids:
                    cam = Camera()
                    C.memmove(C.byref(cam), C.byref(scene.cameras[k]), C.sizeof(cam))
                    if scale != 1:
                        sx, sy = w / float(W), h / float(H)
                        cam.K[0] *= sx; cam.K[2] *= sx; cam.K[4] *= sy; cam.K[5] *= sy
                    cam.width, cam.height = w, h
                    cams.append(cam)
                p.depth_min, p.depth_max = cams[0].depth_min * 0.6, cams[0].depth_max * 1.2

                def rs(a):
                    return a if a.shape[:2] == (h, w) else cv2.resize(a, (w, h), interpolation=cv2.INTER_NEAREST)
                depths = None
                if p.geom_consistency or p.use_APD:
                    depths = [rs(maps[k]["depth"]) for k in ids]
                planes = weak = conf = None
                if p.state != 0:
                    planes = np.concatenate([rs(maps[v]["normal"]), rs(maps[v]["depth"])[..., None]], -1)
                if p.use_APD:
                    weak, conf = rs(maps[v]["weak"]), rs(maps[v]["conf"])
                pl, wk, cf, ms = ref.run_pass([lv_imgs[k] for k in ids], cams, p, planes, weak, conf, depths,
                                              seed=seed_base + 1000 * it + v)
                if v in timed:
                    pm_ms += ms
                depth = pl[..., 3].copy()
                bad = (depth < p.depth_min) | (depth > p.depth_max)
                depth[bad] = 0
                wk[bad] = 2
                maps[v].update(depth=depth, normal=np.ascontiguousarray(pl[..., :3]), weak=wk)
                if p.geom_consistency or p.use_APD:
                    maps[v]["conf"] = cf
                elif "conf" not in maps[v]:
                    maps[v]["conf"] = np.ones((h, w), np.uint8)
            it += 1
    return maps, pm_ms
