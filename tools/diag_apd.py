import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_from_ctx, pull_state, push_state
from apde_mvs_b200.binding import Context, STAGE, default_schedule, default_params
from apde_mvs_b200.scene import make_office_scene
scene = make_office_scene(256, 192, num_views=6, num_src=4, seed=3, weak=0.35, with_color=True)
ctx = Context(0); ctx.load_scene(scene)
sched = default_schedule(); sched.rounds, sched.seed = 2, 5
for p in range(4): ctx.run_schedule_pass(sched, p)
p = default_params(); p.use_APD=1; p.state=1; p.geom_consistency=0; p.weak_peak_radius=6; p.ransac_threshold=0.01-0.00125; p.rotate_time=2; p.max_iterations=3
ctx.problem_setup(1, p, 1, 77)
pb = oracle_from_ctx(ctx, 77, 1)
st = pull_state(ctx)
for f in ("planes", "weak_info", "confidence", "fit_planes"): getattr(pb, f)[...] = st[f]
print("weak hist", np.bincount(st["weak_info"].ravel(), minlength=3), "depth==0:", (st["planes"][...,3]==0).mean())
ctx.problem_stage(STAGE.NEAREST_STRONG); pb.stage("nearest_strong")
ctx.problem_stage(STAGE.GEN_ANCHORS); pb.stage("gen_anchors"); pb.stage("neighbour_update")
push_state(ctx, pb, ("weak_info", "weak_reliable", "anchors"))
planes_before = pb.planes.copy()
ctx.problem_stage(STAGE.INIT); pb.stage("random_init")
st = pull_state(ctx)
d = np.abs(st["costs"] - pb.costs)
for name, m in (("strong", pb.weak_info==1), ("weak", pb.weak_info==0), ("unknown", pb.weak_info==2)):
    print(name, m.sum(), "frac<=1e-3", (d[m]<=1e-3).mean(), "planes close", np.isclose(st["planes"][m], pb.planes[m], rtol=1e-4, atol=1e-5).all(axis=1).mean())
bad = np.argwhere((d>1e-2))
print("bad count", len(bad))
for y,x in bad[:8]:
    print((x,y), "state", pb.weak_info[y,x], "gpu cost", st["costs"][y,x], "orc cost", pb.costs[y,x], "plane gpu", st["planes"][y,x], "plane orc", pb.planes[y,x], "before", planes_before[y,x], "sel gpu %x orc %x"%(st["selected_views"][y,x], pb.selected_views[y,x]))
# per-view costs at a bad pixel
if len(bad):
    y,x = bad[0]
    for v in range(1,5):
        t = np.array([[x,y,v]]); pl = pb.planes[y,x][None]
        mode = 1 if pb.weak_info[y,x]==0 else 0
        print(" view", v, "mode", mode, "gpu", ctx.eval_costs(t, pl, mode), "orc", pb.eval_costs(t, pl, mode), "anchors", pb.anchors[y,x] if mode else "")

# ---- 3-way comparison on this weak-texture scene: product vs CPU oracle vs the reference's own device functions
from oracle import ref_binding as ref
cams, prm = ctx.problem_cameras()
w, h, n = ctx.problem_dims()
imgs = [ctx.problem_image(i) for i in range(n)]
rng = np.random.default_rng(11)
def rep(name, a, b):
    dd = np.abs(a - b)
    print("%-34s max %.3g p99 %.3g p50 %.3g frac<=1e-4 %.5f frac<=1e-3 %.5f" % (name, dd.max(), np.quantile(dd, .99), np.median(dd), (dd <= 1e-4).mean(), (dd <= 1e-3).mean()))
N = 30000
xs, ys, vs = rng.integers(0, w, N), rng.integers(0, h, N), rng.integers(1, n, N)
pl = pb.planes[ys, xs]
t = np.stack([xs, ys, vs], 1)
mine, orac = ctx.eval_costs(t, pl, 0), pb.eval_costs(t, pl, 0)
refc = ref.eval_costs(imgs, [cams[i] for i in range(n)], prm, t, pl, 0)
print("NCC-Old on all pixels of the weak-texture scene (cost hist ref: %s)" % np.histogram(refc, bins=[0, .05, .2, .5, 1, 1.5, 1.999, 2.0])[0])
rep("mine vs reference", mine, refc); rep("oracle vs reference", orac, refc); rep("mine vs oracle", mine, orac)
gx = np.abs(np.diff(imgs[0], axis=1, prepend=0)); lowtex = np.zeros((h, w), bool)
import cv2
std = np.sqrt(np.maximum(cv2.blur(imgs[0] ** 2, (11, 11)) - cv2.blur(imgs[0], (11, 11)) ** 2, 0))
low = std[ys, xs] < 1.5
print("low-texture subset (local std < 1.5 grey levels): %d tuples" % low.sum())
rep("  mine vs reference", mine[low], refc[low]); rep("  oracle vs reference", orac[low], refc[low])
rep("  textured: mine vs reference", mine[~low], refc[~low]); rep("  textured: oracle vs reference", orac[~low], refc[~low])
# NCC-New on weak pixels
wy, wx = np.nonzero(pb.weak_info == 0)
pick = rng.choice(len(wx), min(10000, len(wx)), replace=False)
t = np.stack([wx[pick], wy[pick], rng.integers(1, n, len(pick))], 1)
pl = pb.planes[wy[pick], wx[pick]]
mine, orac = ctx.eval_costs(t, pl, 1), pb.eval_costs(t, pl, 1)
refc = ref.eval_costs(imgs, [cams[i] for i in range(n)], prm, t, pl, 1, weak=pb.weak_info, selected_views=st["selected_views"], anchors=pb.anchors)
print("NCC-New on weak pixels (cost hist ref: %s)" % np.histogram(refc, bins=[0, .05, .2, .5, 1, 1.5, 1.999, 2.0])[0])
rep("mine vs reference", mine, refc); rep("oracle vs reference", orac, refc); rep("mine vs oracle", mine, orac)
