"""Stage-level determinism: replay the first use_APD pass of view 0 twice from identical inputs, compare state after every stage."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from apde_mvs_b200.binding import Context, default_schedule, default_params, STAGE
from apde_mvs_b200.scene import make_office_scene
from helpers import pull_state

scene = make_office_scene(1000, 750, num_views=4, num_src=3, seed=2, arc_deg=15.0)
ctx = Context(0)
sched = default_schedule(); sched.seed = 21
ctx.load_scene(scene)
for p in range(4):
    ctx.run_schedule_pass(sched, p)
p = default_params()
p.geom_factor, p.use_impetus, p.max_iterations = sched.geom_factor, sched.use_impetus, 3
p.use_APD, p.ransac_threshold, p.rotate_time = 1, 0.01 - 0.00125, 2
p.state, p.geom_consistency, p.weak_peak_radius = 1, 0, 6
seq = [("nearest", STAGE.NEAREST_STRONG, ()), ("anchors", STAGE.GEN_ANCHORS, ()), ("init", STAGE.INIT, ())]
for it in range(3):
    seq += [("strong%d.%d" % (it, c), STAGE.PROP_STRONG, (it, c)) for c in (0, 1)]
    seq += [("fit%d" % it, STAGE.RANSAC_FIT, (it,))]
    seq += [("weak%d.%d" % (it, c), STAGE.PROP_WEAK, (it, c)) for c in (0, 1)]
seq += [("depth_normal", STAGE.DEPTH_NORMAL, ()), ("median0", STAGE.MEDIAN, (0, 0)), ("median1", STAGE.MEDIAN, (0, 1)),
        ("to_weak", STAGE.DEPTH_TO_WEAK, ()), ("conf", STAGE.CONFIDENCE, ()), ("refine", STAGE.LOCAL_REFINE, ())]
mode = sys.argv[1] if len(sys.argv) > 1 else "stages"
snaps = []
for rep in range(2):
    ctx.problem_setup(0, p, 1, 1234)
    out = []
    if mode == "stages":
        for name, st, args in seq:
            ctx.problem_stage(st, *args)
            out.append((name, pull_state(ctx)))
    else:
        ctx.problem_run()
        out.append(("run", pull_state(ctx)))
    snaps.append(out)
for (name, a), (_, b) in zip(*snaps):
    msg = []
    for k in a:
        d = a[k] != b[k]
        if a[k].dtype.kind == "f":
            d &= ~(np.isnan(a[k]) & np.isnan(b[k]))
        if d.any():
            msg.append("%s:%d" % (k, d.reshape(d.shape[0], d.shape[1], -1).any(-1).sum()))
    print("%-14s %s" % (name, " ".join(msg) if msg else "identical"))
