"""fp16-texel experiment: (1) are filtered samples bit-identical to the fp32 texture? (2) gather rate under scatter."""
import os, subprocess, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from apde_mvs_b200.binding import Context, default_params
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(1920, 1080, num_views=2, num_src=1, seed=1)
    ctx = Context(0); ctx.load_scene(scene)
    p = default_params(); p.use_APD = 0; p.state = 0
    out = {}
    for scale in (1, 2):
        ctx.problem_setup(0, p, scale, 1)
        w, h, _ = ctx.problem_dims()
        rng = np.random.default_rng(scale)
        xy = np.stack([rng.uniform(-2, w + 2, 200000), rng.uniform(-2, h + 2, 200000)], 1).astype(np.float32)
        out["s%d" % scale] = ctx.debug_tex2d(1, xy)
    ctx.problem_setup(0, p, 1, 1)
    ctx.lib.apde_microbench_pattern.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_double)]
    rates = []
    for spread in (0.0, 2.0, 4.0, 16.0, 64.0):
        g = C.c_double(); ctx.lib.apde_microbench_pattern(ctx._h, 0, spread, C.byref(g)); rates.append(g.value)
    out["rates"] = np.array(rates)
    np.savez(sys.argv[1], **out)
    sys.exit(0)
res = {}
for name, env in (("fp32", "0"), ("fp16", "1")):
    f = "/tmp/fp16_%s.npz" % name
    subprocess.check_call([sys.executable, __file__, f], env=dict(os.environ, APDE_TEX_FP16=env))
    res[name] = np.load(f)
for k in ("s1", "s2"):
    d = np.abs(res["fp32"][k] - res["fp16"][k])
    print("scale %s: fp16 vs fp32 filtered samples: max abs diff %g, bit-identical %.6f" % (k, d.max(), (d == 0).mean()))
print("gather rate Gs/s at spread 0/2/4/16/64 px: fp32 %s | fp16 %s" % (np.round(res["fp32"]["rates"], 0), np.round(res["fp16"]["rates"], 0)))
