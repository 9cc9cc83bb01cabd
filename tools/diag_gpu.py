"""GPU diagnostics (run on the B200 box): texture-filter model probe and 3-way cost comparison
(CUDA product vs CPU oracle vs the reference's own device functions)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_from_ctx, ref_params, to_apde_params  # noqa: E402
from apde_mvs_b200.binding import Context  # noqa: E402
from apde_mvs_b200.scene import make_plane_scene  # noqa: E402
from oracle import binding as orc  # noqa: E402
from oracle import ref_binding as ref  # noqa: E402


def tex_models(img, xy):
    h, w = img.shape
    x, y = xy[:, 0].astype(np.float32), xy[:, 1].astype(np.float32)
    xb, yb = x - np.float32(0.5), y - np.float32(0.5)
    fi, fj = np.floor(xb), np.floor(yb)
    a, b = (xb - fi).astype(np.float64), (yb - fj).astype(np.float64)
    i0 = np.clip(fi.astype(int), 0, w - 1); i1 = np.clip(fi.astype(int) + 1, 0, w - 1)
    j0 = np.clip(fj.astype(int), 0, h - 1); j1 = np.clip(fj.astype(int) + 1, 0, h - 1)
    t00, t10, t01, t11 = img[j0, i0].astype(np.float64), img[j0, i1].astype(np.float64), img[j1, i0].astype(np.float64), img[j1, i1].astype(np.float64)
    out = {}
    for name, q in (("exact", lambda f: f), ("round8", lambda f: np.floor(f * 256 + 0.5) / 256), ("trunc8", lambda f: np.floor(f * 256) / 256),
                    ("round9", lambda f: np.floor(f * 512 + 0.5) / 512)):
        aa, bb = q(a), q(b)
        out[name] = ((1 - aa) * (1 - bb) * t00 + aa * (1 - bb) * t10 + (1 - aa) * bb * t01 + aa * bb * t11).astype(np.float32)
    # fixed-point coordinate model: x*256 rounded to nearest integer first, then -128
    xf = np.floor(x.astype(np.float64) * 256 + 0.5) - 128
    yf = np.floor(y.astype(np.float64) * 256 + 0.5) - 128
    ii, jj = np.floor(xf / 256).astype(int), np.floor(yf / 256).astype(int)
    aa, bb = (xf - ii * 256) / 256, (yf - jj * 256) / 256
    i0 = np.clip(ii, 0, w - 1); i1 = np.clip(ii + 1, 0, w - 1); j0 = np.clip(jj, 0, h - 1); j1 = np.clip(jj + 1, 0, h - 1)
    t00, t10, t01, t11 = img[j0, i0].astype(np.float64), img[j0, i1].astype(np.float64), img[j1, i0].astype(np.float64), img[j1, i1].astype(np.float64)
    out["fixed8"] = ((1 - aa) * (1 - bb) * t00 + aa * (1 - bb) * t10 + (1 - aa) * bb * t01 + aa * bb * t11).astype(np.float32)
    A, B = aa * 256, bb * 256
    W11 = np.floor(A * B / 256 + 0.5)
    out["hw4w"] = (((256 - A - B + W11) * t00 + (A - W11) * t10 + (B - W11) * t01 + W11 * t11) / 256).astype(np.float32)
    return out


def main():
    scene = make_plane_scene(320, 240, num_views=5, num_src=4, seed=1)
    ctx = Context(0)
    ctx.load_scene(scene)
    ctx.problem_setup(2, to_apde_params(ref_params()), 1, 1234)
    rng = np.random.default_rng(0)
    w, h, n = ctx.problem_dims()
    img = ctx.problem_image(1)
    xy = np.stack([rng.uniform(-3, w + 3, 200000), rng.uniform(-3, h + 3, 200000)], 1).astype(np.float32)
    got = ctx.debug_tex2d(1, xy)
    for name, m in tex_models(img, xy).items():
        d = np.abs(got - m)
        print("tex model %-7s max %.6g  mean %.3g  frac exact-equal %.5f" % (name, d.max(), d.mean(), (d == 0).mean()))
    o = np.array([orc.lib().orc_tex2d(img.ctypes.data, w, h, float(x), float(y), 1) for x, y in xy[:5000]], np.float32)
    print("oracle tex_mode1 vs gpu: max %.6g" % np.abs(o - got[:5000]).max())

    # ---- 3-way costs
    pb = oracle_from_ctx(ctx, 1234, 2)
    cams, prm = ctx.problem_cameras()
    imgs = [ctx.problem_image(i) for i in range(n)]
    N = 50000
    xs, ys, vs = rng.integers(0, w, N), rng.integers(0, h, N), rng.integers(1, n, N)
    R, t = scene.Rs[2], scene.ts[2]
    n_w = np.array([-0.15, 0.1, 1.0]); n_w /= np.linalg.norm(n_w)
    n_c = R @ n_w
    n_c = -n_c if n_c[2] > 0 else n_c
    K = scene.K
    d = scene.gt_depth[2][ys, xs]
    X = np.stack([d * (xs - K[0, 2]) / K[0, 0], d * (ys - K[1, 2]) / K[1, 1], d], -1)
    planes = np.concatenate([np.tile(n_c, (N, 1)), (-(X @ n_c))[:, None]], 1).astype(np.float32)
    k = N // 2
    planes[:k, 3] *= rng.uniform(0.95, 1.05, k).astype(np.float32)
    planes[:k, :2] += rng.normal(0, 0.05, (k, 2)).astype(np.float32)
    tuples = np.stack([xs, ys, vs], 1)
    mine = ctx.eval_costs(tuples, planes, 0)
    orac = pb.eval_costs(tuples, planes, 0)
    pb0 = oracle_from_ctx(ctx, 1234, 2, tex_mode=0)
    orac0 = pb0.eval_costs(tuples, planes, 0)
    refc = ref.eval_costs(imgs, [cams[i] for i in range(n)], prm, tuples, planes, 0)

    def rep(name, a, b):
        dd = np.abs(a - b)
        print("%-28s max %.3g  p99 %.3g  p50 %.3g  frac<=1e-4 %.5f  frac<=2e-4 %.5f" % (name, dd.max(), np.quantile(dd, 0.99), np.median(dd), (dd <= 1e-4).mean(), (dd <= 2e-4).mean()))
    rep("mine vs reference", mine, refc)
    rep("oracle(8bit) vs reference", orac, refc)
    rep("oracle(exact) vs reference", orac0, refc)
    rep("mine vs oracle(8bit)", mine, orac)
    print("cost histogram (reference):", np.histogram(refc, bins=[0, 0.05, 0.2, 0.5, 1.0, 1.5, 1.999, 2.0])[0])
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_costs.npz"), mine=mine, orac=orac, refc=refc, tuples=tuples, planes=planes)


if __name__ == "__main__":
    main()
