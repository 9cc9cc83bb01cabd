"""diagnostic: outliers of the segment-label cost path (ours / oracle / reference device function)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_from_ctx, pull_state, push_state
from test_gpu_sa_mask import make_labels, _setup, _ref_costs, SEED
from apde_mvs_b200.binding import Context, FIELD, STAGE, default_schedule
from apde_mvs_b200.scene import make_office_scene

ctx = Context(0)
scene = make_office_scene(256, 192, num_views=6, num_src=4, seed=3, weak=0.35, with_color=True)
ctx.load_scene(scene)
sched = default_schedule(); sched.rounds, sched.seed = 2, 5
for pidx in range(4):
    ctx.run_schedule_pass(sched, pidx)
lab = make_labels(256, 192, 7)
_setup(ctx, sched, 2, lab)
pb = oracle_from_ctx(ctx, SEED, 2)
st = pull_state(ctx)
for f in ("planes", "weak_info", "confidence"):
    getattr(pb, f)[...] = st[f]
cams, _ = ctx.problem_cameras()
w, h, n = ctx.problem_dims()
rng = np.random.default_rng(0)
m = 20000
xs, ys, vs = rng.integers(0, w, m), rng.integers(0, h, m), rng.integers(1, n, m)
R = np.array(cams[0].R, np.float32).reshape(3, 3)
pw = st["planes"][ys, xs]
nc = pw[:, :3] @ R.T
depth = np.where(pw[:, 3] > 0, pw[:, 3], 4.0)
K = np.array(cams[0].K, np.float32).reshape(3, 3)
X = np.stack([depth * (xs - K[0, 2]) / K[0, 0], depth * (ys - K[1, 2]) / K[1, 1], depth], 1)
bad = np.linalg.norm(nc, axis=1) < 0.5
nc[bad] = [0, 0, -1]
planes = np.concatenate([nc, -(X * nc).sum(1, keepdims=True)], 1).astype(np.float32)
k = m // 2
planes[:k, 3] *= rng.uniform(0.95, 1.05, k).astype(np.float32)
tuples = np.stack([xs, ys, vs], 1)
got = ctx.eval_costs(tuples, planes, 0)
pb.set_sa_mask(ctx.problem_get(FIELD.SA_MASK))
want = pb.eval_costs(tuples, planes, 0)
rc = _ref_costs(ctx, pb, lab, tuples, planes, 0)
for name, a, b in (("ours-oracle", got, want), ("ours-ref", got, rc), ("oracle-ref", want, rc)):
    d = np.abs(a - b)
    print("%s: frac<=1e-4 %.5f <=1e-3 %.5f max %.3g n>1e-2: %d" % (name, (d <= 1e-4).mean(), (d <= 1e-3).mean(), d.max(), (d > 1e-2).sum()))


def tap_count(x, y):
    sign = [1, 1, -1, -1, 1, -1, -1, 1]
    off = [1, 1, 3, 1, 1, 3, 1, 5, 3, 3, 5, 1, 5, 3, 3, 5, 5, 5]
    c = 0
    for i in range(4):
        for j in range(9):
            rx, ry = x + off[2 * j] * sign[2 * i], y + off[2 * j + 1] * sign[2 * i + 1]
            if rx < 0 or rx >= w or ry < 0 or ry >= h:
                continue
            if lab[ry, rx] != lab[y, x]:
                break
            c += 1
    return c


d = np.abs(got - rc)
worst = np.argsort(-d)[:25]
img = ctx.problem_image(0)
for i in worst:
    x, y = xs[i], ys[i]
    patch = img[max(y - 5, 0):y + 6, max(x - 5, 0):x + 6]
    print("px (%3d,%3d) v%d label %3d taps %2d  ours %.5f oracle %.5f ref %.5f  patch std %.3f" % (x, y, vs[i], lab[y, x], tap_count(x, y), got[i], want[i], rc[i], patch.std()))
cnt = np.array([tap_count(xs[i], ys[i]) for i in range(m)])
for c in range(0, 37):
    sel = cnt == c
    if sel.sum():
        print("taps %2d: n %5d  ours-ref frac<=1e-4 %.4f  oracle-ref %.4f" % (c, sel.sum(), (np.abs(got - rc)[sel] <= 1e-4).mean(), (np.abs(want - rc)[sel] <= 1e-4).mean()))
