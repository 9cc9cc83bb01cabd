"""8-bit UNORM texel path: bit-exactness of fetch() against the float32 texture, gather rate, end-to-end effect."""
import os, subprocess, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from apde_mvs_b200.binding import Context, default_params
    from apde_mvs_b200.scene import make_office_scene
    scene = make_office_scene(1920, 1080, num_views=2, num_src=1, seed=1)
    ctx = Context(0); ctx.load_scene(scene)
    p = default_params(); p.use_APD = 0; p.state = 0
    out = {}
    for scale in (4, 2, 1):
        ctx.problem_setup(0, p, scale, 1)
        w, h, _ = ctx.problem_dims()
        rng = np.random.default_rng(scale)
        xy = np.stack([rng.uniform(-2, w + 2, 1000000), rng.uniform(-2, h + 2, 1000000)], 1).astype(np.float32)
        out["s%d" % scale] = ctx.debug_tex2d(1, xy)
    ctx.lib.apde_microbench_pattern.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_double)]
    rates = []
    for spread in (0.0, 2.0, 4.0, 16.0, 64.0):
        g = C.c_double(); ctx.lib.apde_microbench_pattern(ctx._h, 0, spread, C.byref(g)); rates.append(g.value)
    out["rates"] = np.array(rates)
    np.savez(sys.argv[1], **out)
    sys.exit(0)
res = {}
for name, env in (("fp32", "0"), ("u8", "1")):
    f = "/tmp/u8_%s.npz" % name
    subprocess.check_call([sys.executable, __file__, f], env=dict(os.environ, APDE_TEX_U8=env))
    res[name] = np.load(f)
for k in ("s1", "s2", "s4"):
    d = np.abs(res["fp32"][k] - res["u8"][k])
    print("scale %s: UNORM texels vs fp32 texels, filtered samples (1e6 probes): max abs diff %g, bit-identical %.7f" % (k[1:], d.max(), (d == 0).mean()))
print("raw gather rate Gs/s at spread 0/2/4/16/64 px: fp32 %s | u8 %s" % (np.round(res["fp32"]["rates"], 0), np.round(res["u8"]["rates"], 0)))
