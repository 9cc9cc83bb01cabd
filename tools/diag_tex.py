"""1-D probe of the texture unit's interpolation weight quantisation (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import ref_params, to_apde_params
from apde_mvs_b200.binding import Context
from apde_mvs_b200.scene import make_plane_scene
scene = make_plane_scene(320, 240, num_views=5, num_src=4, seed=1)
ctx = Context(0); ctx.load_scene(scene)
ctx.problem_setup(2, to_apde_params(ref_params()), 1, 1)
img = ctx.problem_image(1)
h, w = img.shape
# find strong horizontal gradients
g = np.abs(img[:, 1:] - img[:, :-1])
js, is_ = np.unravel_index(np.argsort(g.ravel())[-8:], g.shape)
K = 8192
res = {}
for j, i in zip(js, is_):
    fr = np.arange(K, dtype=np.float64) / K
    x = (i + 0.5 + fr).astype(np.float32)
    y = np.full(K, j + 0.5, np.float32)
    got = ctx.debug_tex2d(1, np.stack([x, y], 1))
    t0, t1 = float(img[j, i]), float(img[j, i + 1])
    alpha = (got.astype(np.float64) - t0) / (t1 - t0)
    res["a_%d_%d" % (j, i)] = alpha
    res["x_%d_%d" % (j, i)] = x
    lv = np.unique(np.round(alpha * 256, 3))
    print("pixel (%d,%d) T0=%g T1=%g: %d distinct alpha levels; first 6 (x256): %s" % (j, i, t0, t1, len(lv), lv[:6]))
    # where do the steps happen, in units of 1/256 of frac?
    steps = np.nonzero(np.diff(np.round(alpha * 256)) != 0)[0]
    print("   step positions (frac*256) first 6:", np.round((steps[:6] + 1) / K * 256, 4))
# same in y
res["img"] = img
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_tex.npz"), **res)
# large-coordinate behaviour: same fractional probe at i ~ 300
# ---- 2-D probe: random coordinates, saved for offline model fitting
rng = np.random.default_rng(0)
xy = np.stack([rng.uniform(-3, w + 3, 100000), rng.uniform(-3, h + 3, 100000)], 1).astype(np.float32)
got = ctx.debug_tex2d(1, xy)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_tex2d.npz"), img=img, xy=xy, got=got)
