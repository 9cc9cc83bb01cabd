"""One strong half-sweep from identical state with both kernels (APDE_STRONG_V1 toggled in-process); report differences."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from apde_mvs_b200.binding import Context, default_params, STAGE, FIELD
from apde_mvs_b200.scene import make_office_scene
from helpers import pull_state

scene = make_office_scene(320, 240, num_views=5, num_src=4, seed=2, arc_deg=15.0)
ctx = Context(0)
ctx.load_scene(scene)
p = default_params(); p.use_APD = 0; p.state = 0
ctx.problem_setup(2, p, 1, 5)
ctx.problem_stage(STAGE.INIT)
s0 = pull_state(ctx)
res = {}
for v1 in ("1", "0"):
    os.environ["APDE_STRONG_V1"] = v1
    for k, f in (("planes", FIELD.PLANES), ("costs", FIELD.COSTS), ("selected_views", FIELD.SELECTED_VIEWS), ("view_weight", FIELD.VIEW_WEIGHT)):
        ctx.problem_set(f, s0[k])
    ctx.problem_stage(STAGE.PROP_STRONG, 0, 0)
    res[v1] = pull_state(ctx)
a, b = res["1"], res["0"]
for k in ("planes", "costs", "selected_views", "view_weight"):
    d = a[k] != b[k]
    if d.ndim == 3: d = d.any(-1)
    print(k, "differ:", int(d.sum()), "of", d.size)
d = (a["planes"] != b["planes"]).any(-1)
ys, xs = np.nonzero(d)
for y, x in list(zip(ys, xs))[:6]:
    print((x, y), "v1", a["planes"][y, x], a["costs"][y, x], "new", b["planes"][y, x], b["costs"][y, x], "sel", bin(a["selected_views"][y, x]), bin(b["selected_views"][y, x]))
print("changed by v1:", int((a["planes"] != s0["planes"]).any(-1).sum()), " by new:", int((b["planes"] != s0["planes"]).any(-1).sum()))
d = a["costs"] != b["costs"]
ys, xs = np.nonzero(d)
dd = np.abs(a["costs"] - b["costs"])[d]
print("cost diffs: max %.3g median %.3g" % (dd.max(), np.median(dd)))
for y, x in list(zip(ys, xs))[:10]:
    print((x, y), "v1 cost %.9g new cost %.9g  plane changed from init: %s  popc(sel) %d" % (a["costs"][y, x], b["costs"][y, x],
          bool((a["planes"][y, x] != s0["planes"][y, x]).any()), bin(a["selected_views"][y, x]).count("1")))
