import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apde_mvs_b200.binding import Context, default_params
from apde_mvs_b200.scene import make_plane_scene
scene = make_plane_scene(1920, 1080, num_views=2, num_src=1, seed=1)
ctx = Context(0); ctx.load_scene(scene)
p = default_params(); p.use_APD = 0; p.state = 0
ctx.problem_setup(0, p, 1, 1)
ctx.lib.apde_microbench_pattern.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_double)]
print("fp32 / tex peak:", ctx.microbench())
for mode in (0, 1):
    for spread in (0.0, 4.0, 16.0, 64.0, 256.0):
        g = C.c_double()
        rc = ctx.lib.apde_microbench_pattern(ctx._h, mode, spread, C.byref(g))
        print("mode %d (%s) spread %5.0f px: %.1f Gsamples/s" % (mode, "thread/eval" if mode == 0 else "quad/eval", spread, g.value), rc)
