"""A/B: thread-per-evaluation kernels (APDE_THREAD_KERNELS=1) vs the quad engine: results must be bit-identical."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from apde_mvs_b200.binding import Context, default_schedule
    from apde_mvs_b200.scene import make_office_scene
    scene = make_office_scene(384, 256, num_views=6, num_src=5, seed=3, weak=0.3)
    ctx = Context(0); ctx.load_scene(scene)
    s = default_schedule(); s.rounds, s.seed = 2, 9
    t = ctx.run_schedule(s)
    out = {}
    for v in range(6):
        d, n, w, c = ctx.view_download(v)
        out["d%d" % v], out["n%d" % v], out["w%d" % v], out["c%d" % v] = d, n, w, c
    np.savez(sys.argv[2], ms=t.patchmatch_ms, evals=np.array([t.evals_ncc_old, t.evals_ncc_new, t.evals_geom]), **out)
    sys.exit(0)
res = {}
for name, env in (("thread", "0"), ("quad", "1")):
    f = "/tmp/ab_%s.npz" % name
    subprocess.check_call([sys.executable, __file__, "child", f], env=dict(os.environ, APDE_QUAD_KERNELS=env))
    res[name] = np.load(f)
a, b = res["thread"], res["quad"]
print("patchmatch ms: thread %.1f quad %.1f ; evals %s vs %s" % (a["ms"], b["ms"], a["evals"], b["evals"]))
for k in a.files:
    if k in ("ms", "evals"): continue
    same = np.array_equal(a[k], b[k], equal_nan=True)
    if not same:
        print(k, "differs: frac equal %.6f" % (np.isclose(a[k], b[k], equal_nan=True).mean()))
print("bit-identical:", all(np.array_equal(a[k], b[k], equal_nan=True) for k in a.files if k not in ("ms", "evals")))
