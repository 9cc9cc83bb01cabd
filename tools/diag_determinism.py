"""Run the same schedule twice with the same seed and report the first pass / map whose bits differ."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apde_mvs_b200.binding import Context, default_schedule
from apde_mvs_b200.scene import make_office_scene

W, H, V, N = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1000, 750, 4, 3)))
weak = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
scene = make_office_scene(W, H, num_views=V, num_src=N, seed=2, arc_deg=15.0, weak=weak)
ctx = Context(0)
sched = default_schedule()
sched.seed = 21
sched.jacobi = int(os.environ.get("JACOBI", "0"))
if os.environ.get("ROUNDS"): sched.rounds = int(os.environ["ROUNDS"])
runs = []
for rep in range(2):
    ctx.load_scene(scene)
    npass = ctx.num_passes(sched)
    snaps = []
    for p in range(npass):
        ctx.run_schedule_pass(sched, p)
        snaps.append([ctx.view_download(v) for v in range(V)])
    runs.append(snaps)
names = ("depth", "normal", "weak", "conf")
for p in range(len(runs[0])):
    for v in range(V):
        for k in range(4):
            a, b = runs[0][p][v][k], runs[1][p][v][k]
            if a.shape != b.shape:
                print("pass %d view %d %s: shapes differ" % (p, v, names[k])); continue
            d = (a != b)
            if d.ndim == 3:
                d = d.any(-1)
            if d.any():
                ys, xs = np.nonzero(d)
                print("pass %d view %d %-6s: %7d px differ (%.5f); first at (%d,%d): %s / %s; weak there %s / %s" % (
                    p, v, names[k], d.sum(), d.mean(), xs[0], ys[0], a[ys[0], xs[0]], b[ys[0], xs[0]],
                    runs[0][p][v][2][ys[0], xs[0]], runs[1][p][v][2][ys[0], xs[0]]))
print("done")
