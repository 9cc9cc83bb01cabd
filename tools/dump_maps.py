"""Run the whole schedule on a synthetic office scene and dump all maps to an .npz (for A/B comparisons of kernel variants
selected by environment variables, which the library reads once per process).
usage: [APDE_DUMP_SA=1] dump_maps.py out.npz [W H V N weak rounds]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apde_mvs_b200.binding import Context, default_schedule
from apde_mvs_b200.scene import make_label_map, make_office_scene

out = sys.argv[1]
W, H, V, N = (int(x) for x in (sys.argv[2:6] if len(sys.argv) >= 6 else (640, 480, 5, 4)))
weak = float(sys.argv[6]) if len(sys.argv) > 6 else 0.2
rounds = int(sys.argv[7]) if len(sys.argv) > 7 else 2
scene = make_office_scene(W, H, num_views=V, num_src=N, seed=2, arc_deg=15.0, weak=weak)
ctx = Context(0)
ctx.load_scene(scene)
if os.environ.get("APDE_DUMP_SA") == "1":  # segment-label maps on every view (half size: the nearest resize runs too)
    for v in range(V):
        ctx.view_set_sa_mask(v, make_label_map(W // 2, H // 2, 40 + v, zero_share=0.2))
sched = default_schedule()
sched.seed, sched.rounds = 21, rounds
t = ctx.run_schedule(sched)
maps = {}
for v in range(V):
    d, n, w, c = ctx.view_download(v)
    maps["d%d" % v], maps["n%d" % v], maps["w%d" % v], maps["c%d" % v] = d, n, w, c
np.savez(out, **maps)
print("patchmatch %.1f ms, evals %d/%d/%d" % (t.patchmatch_ms, t.evals_ncc_old, t.evals_ncc_new, t.evals_geom))
