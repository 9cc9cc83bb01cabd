"""Design input for a shared-memory sampling path of the DepthToWeak sweep.

For the default bench geometry (office scene, 11 views on a 60 degree arc, 10 source views) this computes, per CTA-sized
tile of reference pixels and per source view, the bounding box in the source image of EVERY sample the sweep takes: 61
hypotheses `fb / (disp + k)`, k = -30..30 (APD.cu:2157-2165), times the 36 taps of the 11x11 patch (stride 2), warped by
the plane-induced homography of the pixel's own plane (ground-truth depth, normal from the depth gradient).  CPU only
(numpy); no image content is needed, only cameras and depth.

    python tools/analyze_sweep_footprint.py [W H]        (default 960 540: the geometry scales with the image)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apde_mvs_b200.scene import make_office_scene  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
V, N = 11, 10
scene = make_office_scene(W, H, V, N, seed=2)
K = scene.K
Ki = np.linalg.inv(K)
ref = 5
R0, t0 = scene.Rs[ref], scene.ts[ref]
C0 = -R0.T @ t0
gt = scene.gt_depth[ref].astype(np.float64)
ys, xs = np.mgrid[0:H, 0:W]
# camera-frame points and normals (from depth gradients; facets are planar)
X = np.stack([(xs - K[0, 2]) / K[0, 0] * gt, (ys - K[1, 2]) / K[1, 1] * gt, gt], -1)
dx = np.gradient(X, axis=1)
dy = np.gradient(X, axis=0)
nrm = np.cross(dx, dy)
nrm /= np.linalg.norm(nrm, axis=2, keepdims=True) + 1e-12
nrm[nrm[..., 2] > 0] *= -1
dist = -(nrm * X).sum(-1)  # plane: n.X + dist = 0

srcs = scene.pairs[ref]
base = []
for s in srcs:
    Cs = -scene.Rs[s].T @ scene.ts[s]
    base.append(np.linalg.norm(Cs - C0))
mean_base = float(np.mean(base))  # all views selected: the worst case for the step size
taps = np.array([-5, -3, -1, 1, 3, 5], np.float64)
ks = np.arange(-30, 31, dtype=np.float64)


def tile_boxes(tw, th):
    out = []
    for ty in range(8, H - th - 8, max(th, H // 12)):
        for tx in range(8, W - tw - 8, max(tw, W // 12)):
            sl = (slice(ty, ty + th), slice(tx, tx + tw))
            if (gt[sl] <= 0).any():
                continue
            px, py = xs[sl].ravel().astype(np.float64), ys[sl].ravel().astype(np.float64)
            n, d0 = nrm[sl].reshape(-1, 3), gt[sl].ravel()
            fb = K[0, 0] * mean_base
            disp = fb / d0
            for s in srcs:
                Rs, ts = scene.Rs[s], scene.ts[s]
                Rrel = Rs @ R0.T
                trel = ts - Rrel @ t0
                lo, hi = np.array([1e9, 1e9]), np.array([-1e9, -1e9])
                for k in ks:
                    dep = fb / (disp + k)
                    ok = (dep > 0.6 * 2.0) & (dep < 1.2 * 12.0)
                    if not ok.any():
                        continue
                    # plane through the pixel at depth dep with the pixel's normal: distance w = -n.X(dep)
                    Xc = np.stack([(px - K[0, 2]) / K[0, 0] * dep, (py - K[1, 2]) / K[1, 1] * dep, dep], 1)
                    w = -(n * Xc).sum(1)
                    for i in taps:
                        for j in taps:
                            q = np.stack([px + i, py + j, np.ones_like(px)], 1) @ Ki.T
                            zq = -w / (n * q).sum(1)  # depth of the tap's ray on the plane
                            P = q * zq[:, None]
                            S = (P @ Rrel.T + trel) @ K.T
                            uv = S[:, :2] / S[:, 2:3]
                            uv = uv[ok & np.isfinite(uv).all(1)]
                            if len(uv):
                                lo, hi = np.minimum(lo, uv.min(0)), np.maximum(hi, uv.max(0))
                if hi[0] > lo[0]:
                    out.append((hi[0] - lo[0] + 2, hi[1] - lo[1] + 2))
    return np.array(out)


for tw, th in ((128, 1), (32, 4), (16, 8)):
    b = tile_boxes(tw, th)
    area = b[:, 0] * b[:, 1]
    print("tile %3dx%d: %4d (tile, view) boxes; width  median %5.0f p90 %5.0f max %5.0f | height median %4.0f p90 %4.0f max %4.0f | "
          "texels median %6.0f p90 %6.0f max %7.0f" % (tw, th, len(b), np.median(b[:, 0]), np.quantile(b[:, 0], .9), b[:, 0].max(),
                                                       np.median(b[:, 1]), np.quantile(b[:, 1], .9), b[:, 1].max(),
                                                       np.median(area), np.quantile(area, .9), area.max()))
print("(%dx%d, %d source views, mean baseline %.3f m; samples per (tile, view): 128 px x 61 depths x 36 taps = %d)" % (W, H, N, mean_base, 128 * 61 * 36))
