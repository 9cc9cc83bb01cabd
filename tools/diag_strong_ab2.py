"""Per pass: run the pass with k_prop_strong_v1 and with the compacted kernel from the SAME maps; report differing pixels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apde_mvs_b200.binding import Context, default_schedule
from apde_mvs_b200.scene import make_office_scene
V = 4
scene = make_office_scene(320, 240, num_views=V, num_src=3, seed=2, arc_deg=15.0, weak=0.25)
ctx = Context(0)
ctx.load_scene(scene)
sched = default_schedule(); sched.seed, sched.rounds = 21, 2
def snap(): return [ctx.view_download(v) for v in range(V)]
def restore(m):
    for v in range(V): ctx.view_upload(v, *m[v])
def eq(x, y):
    return np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
for p in range(ctx.num_passes(sched)):
    before = snap() if p > 0 else None
    os.environ["APDE_STRONG_V1"] = "1"
    ctx.run_schedule_pass(sched, p)
    a = snap()
    if p == 0:
        ctx.load_scene(scene)
    else:
        restore(before)
    os.environ["APDE_STRONG_V1"] = "0"
    ctx.run_schedule_pass(sched, p)
    b = snap()
    msg = []
    for v in range(V):
        for k, name in enumerate(("depth", "normal", "weak", "conf")):
            if not eq(a[v][k], b[v][k]):
                d = a[v][k] != b[v][k]
                if d.ndim == 3: d = d.any(-1)
                msg.append("view %d %s: %d px" % (v, name, int(d.sum())))
    print("pass %d: %s" % (p, "; ".join(msg) if msg else "identical"))
    # continue from the v1 result so that later passes start from the same maps in both variants
    restore(a)
