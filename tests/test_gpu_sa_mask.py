"""GPU parity tests of the segment-label (SAM mask) path: NCC-Old branch B (APD.cu:664-719), the label tests of NCC-New
(APD.cu:493-497, 526-530) and the stages that call them, against the CPU oracle AND against the reference's own device
functions / RunPatchMatch (oracle/_ref) on identical inputs.  The label maps are synthetic (running SAM is out of scope);
they have the file format and value range of tools/run_SAM.py (uint8, 0 = no segment)."""
import numpy as np
import pytest

from helpers import oracle_from_ctx, plane_depth, pull_state, push_state
from test_gpu_pipeline import _params_for_pass

pytestmark = pytest.mark.gpu
SEED = 91


from apde_mvs_b200.scene import make_label_map as make_labels  # noqa: E402


@pytest.fixture(scope="module")
def office_sa(ctx):
    from apde_mvs_b200.scene import make_office_scene
    from apde_mvs_b200.binding import default_schedule
    scene = make_office_scene(256, 192, num_views=6, num_src=4, seed=3, weak=0.35, with_color=True)
    ctx.load_scene(scene)
    sched = default_schedule()
    sched.rounds, sched.seed = 2, 5
    for pidx in range(0, 4):  # round 0 (half resolution) never reads a label map (APD.cpp:614)
        ctx.run_schedule_pass(sched, pidx)
    return scene, sched


def _setup(ctx, sched, view, labels, geom=False):
    p, scale = _params_for_pass(ctx, sched, 5 if geom else 4)
    assert p.use_APD == 1 and scale == 1
    p.use_sa = 1
    ctx.view_set_sa_mask(view, labels)
    ctx.problem_setup(view, p, scale, SEED)
    return p


def test_label_map_resize_and_switch(ctx, office_sa):
    """the file's map is nearest-resized to the working size as cv::resize(INTER_NEAREST) does (APD.cpp:645-648); without a
    map, with use_sa = 0 or in a pass without use_APD the problem has none"""
    import cv2
    from apde_mvs_b200.binding import FIELD, ApdeError
    scene, sched = office_sa
    for shape in ((192, 256), (96, 128), (150, 201), (384, 512)):
        lab = make_labels(shape[1], shape[0], 3)
        _setup(ctx, sched, 1, lab)
        got = ctx.problem_get(FIELD.SA_MASK)
        want = lab if shape == (192, 256) else cv2.resize(lab, (256, 192), interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(got, want), shape
    p, scale = _params_for_pass(ctx, sched, 4)
    p.use_sa = 0
    ctx.problem_setup(1, p, scale, SEED)
    with pytest.raises(ApdeError):
        ctx.problem_get(FIELD.SA_MASK)
    ctx.view_set_sa_mask(1, None)
    p.use_sa = 1
    ctx.problem_setup(1, p, scale, SEED)
    with pytest.raises(ApdeError):
        ctx.problem_get(FIELD.SA_MASK)


def _ref_costs(ctx, pb, lab, tuples, planes, mode):
    from oracle import ref_binding as ref
    if not ref.available():
        return None
    _, _, n = ctx.problem_dims()
    cams, prm = ctx.problem_cameras()
    imgs = [ctx.problem_image(i) for i in range(n)]
    ref.set_sa_mask(lab)
    try:
        return ref.eval_costs(imgs, [cams[i] for i in range(n)], prm, tuples, planes, mode, depths=pb.depths, weak=pb.weak_info,
                              selected_views=pb.selected_views, anchors=pb.anchors)
    finally:
        ref.set_sa_mask(None)


def test_ncc_old_with_labels(ctx, office_sa):
    """branch B is taken where the label at the PROJECTED point's index is non-zero (quirk of APD.cu:619-621): ours == oracle
    == the reference's own ComputeBilateralNCCOld, and the label map must matter"""
    from apde_mvs_b200.binding import FIELD
    scene, sched = office_sa
    lab = make_labels(256, 192, 7)
    _setup(ctx, sched, 2, lab)
    pb = oracle_from_ctx(ctx, SEED, 2)
    st = pull_state(ctx)
    for f in ("planes", "weak_info", "confidence"):
        getattr(pb, f)[...] = st[f]
    cams, _ = ctx.problem_cameras()
    w, h, n = ctx.problem_dims()
    rng = np.random.default_rng(0)
    m = 20000
    xs, ys, vs = rng.integers(0, w, m), rng.integers(0, h, m), rng.integers(1, n, m)
    # hypotheses: the current maps' planes in the camera frame (mostly converged after round 0), half of them perturbed
    R = np.array(cams[0].R, np.float32).reshape(3, 3)
    pw = st["planes"][ys, xs]
    nc = pw[:, :3] @ R.T
    depth = np.where(pw[:, 3] > 0, pw[:, 3], 4.0)
    K = np.array(cams[0].K, np.float32).reshape(3, 3)
    X = np.stack([depth * (xs - K[0, 2]) / K[0, 0], depth * (ys - K[1, 2]) / K[1, 1], depth], 1)
    bad = np.linalg.norm(nc, axis=1) < 0.5
    nc[bad] = [0, 0, -1]
    planes = np.concatenate([nc, -(X * nc).sum(1, keepdims=True)], 1).astype(np.float32)
    k = m // 2
    planes[:k, 3] *= rng.uniform(0.95, 1.05, k).astype(np.float32)
    tuples = np.stack([xs, ys, vs], 1)
    got = ctx.eval_costs(tuples, planes, 0)
    plain = pb.eval_costs(tuples, planes, 0)          # oracle without the map
    pb.set_sa_mask(ctx.problem_get(FIELD.SA_MASK))
    want = pb.eval_costs(tuples, planes, 0)
    d = np.abs(got - want)
    changed = np.abs(want - plain) > 1e-3
    print("ncc_old + labels vs oracle: max %.3g p99 %.3g frac<=1e-4 %.5f; the map changes %.3f of the costs" % (
        d.max(), np.quantile(d, 0.99), (d <= 1e-4).mean(), changed.mean()))
    assert changed.mean() > 0.2
    assert (d <= 1e-4).mean() >= 0.997 and (d <= 1e-3).mean() >= 0.998  # measured 0.99770: oracle = correctly rounded MUFU
    rc = _ref_costs(ctx, pb, lab, tuples, planes, 0)
    if rc is not None:
        dr, dro = np.abs(got - rc), np.abs(want - rc)
        print("  vs the reference's device function: ours frac<=1e-4 %.5f (max %.3g); oracle frac<=1e-4 %.5f" % (
            (dr <= 1e-4).mean(), dr.max(), (dro <= 1e-4).mean()))
        assert np.array_equal(got, rc)  # the product equals the reference's own ComputeBilateralNCCOld bit for bit
        assert (dro <= 1e-4).mean() >= 0.997
    ctx.view_set_sa_mask(2, None)


def test_apd_stages_with_labels(ctx, office_sa):
    """anchors -> init (NCC-New with label tests on WEAK pixels, NCC-Old B elsewhere) -> strong / weak propagation ->
    DepthToWeak -> LocalRefine, stage by stage from identical state against the oracle; deformable costs also against the
    reference's own ComputeBilateralNCCNew"""
    from apde_mvs_b200.binding import FIELD, STAGE
    scene, sched = office_sa
    lab = make_labels(256, 192, 11, zero_share=0.15)
    _setup(ctx, sched, 1, lab)
    pb = oracle_from_ctx(ctx, SEED, 1)
    pb.set_sa_mask(ctx.problem_get(FIELD.SA_MASK))
    st = pull_state(ctx)
    for f in ("planes", "weak_info", "confidence", "fit_planes"):
        getattr(pb, f)[...] = st[f]
    ctx.problem_stage(STAGE.NEAREST_STRONG)
    pb.stage("nearest_strong")
    ctx.problem_stage(STAGE.GEN_ANCHORS)
    pb.stage("gen_anchors")
    pb.stage("neighbour_update")
    push_state(ctx, pb, ("weak_info", "weak_reliable", "anchors"))
    cams, _ = ctx.problem_cameras()
    ctx.problem_stage(STAGE.INIT)
    pb.stage("random_init")
    st = pull_state(ctx)
    d = np.abs(st["costs"] - pb.costs)
    wk = pb.weak_info == 0
    print("init with labels (%d weak px): |dcost|<=1e-3 weak %.5f all %.5f; masks equal %.5f" % (
        wk.sum(), (d[wk] <= 1e-3).mean(), (d <= 1e-3).mean(), (st["selected_views"] == pb.selected_views).mean()))
    assert wk.sum() > 500
    assert (d[wk] <= 1e-3).mean() >= 0.999 and (d <= 1e-3).mean() >= 0.999  # measured 0.99972 / 0.99969
    push_state(ctx, pb, ("planes", "costs", "selected_views"))
    # deformable cost on the weak pixels
    ys, xs = np.nonzero(wk)
    rng = np.random.default_rng(5)
    pick = rng.choice(len(xs), min(8000, len(xs)), replace=False)
    _, _, n = ctx.problem_dims()
    tuples = np.stack([xs[pick], ys[pick], rng.integers(1, n, len(pick))], 1)
    planes = pb.planes[ys[pick], xs[pick]]
    got, want = ctx.eval_costs(tuples, planes, 1), pb.eval_costs(tuples, planes, 1)
    keep = pb.sa_mask
    pb.set_sa_mask(None)
    plain = pb.eval_costs(tuples, planes, 1)
    pb.set_sa_mask(keep)
    dd = np.abs(got - want)
    print("ncc_new + labels: max %.3g p99 %.3g frac<=1e-4 %.5f; the map changes %.3f of the costs" % (
        dd.max(), np.quantile(dd, 0.99), (dd <= 1e-4).mean(), (np.abs(want - plain) > 1e-3).mean()))
    assert (np.abs(want - plain) > 1e-3).mean() > 0.05
    assert (dd <= 1e-4).mean() >= 0.998 and (dd <= 1e-3).mean() >= 0.998  # measured 0.99859
    rc = _ref_costs(ctx, pb, keep, tuples, planes, 1)
    if rc is not None:
        dr = np.abs(got - rc)
        print("  vs the reference's device function: frac<=1e-4 %.5f max %.3g" % ((dr <= 1e-4).mean(), dr.max()))
        assert np.array_equal(got, rc)  # bit for bit the reference's own ComputeBilateralNCCNew

    def depth_agreement(sel):
        s = pull_state(ctx)
        dg, do = plane_depth(s["planes"], cams[0]), plane_depth(pb.planes, cams[0])
        with np.errstate(all="ignore"):
            ok = (np.abs(dg - do) <= 0.01 * np.abs(do)) | (~np.isfinite(do) & ~np.isfinite(dg))
        return float(ok[sel].mean())

    for color in (0, 1):
        ctx.problem_stage(STAGE.PROP_STRONG, 0, color)
        pb.stage("propagate_strong", 0, color)
        frac = depth_agreement(~wk)
        print("strong propagation colour %d: depth within 1%% of oracle %.5f" % (color, frac))
        assert frac >= 0.9995  # measured 1.0
        push_state(ctx, pb, ("planes", "costs", "selected_views", "view_weight"))
    ctx.problem_stage(STAGE.RANSAC_FIT, 0)
    pb.stage("ransac_fit", 0)
    push_state(ctx, pb, ("fit_planes",))
    for color in (0, 1):
        ctx.problem_stage(STAGE.PROP_WEAK, 0, color)
        pb.stage("propagate_weak", 0, color)
        frac = depth_agreement(wk)
        print("weak propagation colour %d: weak-pixel depth within 1%% of oracle %.5f" % (color, frac))
        assert frac >= 0.999  # measured 1.0
        push_state(ctx, pb, ("planes", "costs", "selected_views", "view_weight"))
    ctx.problem_stage(STAGE.DEPTH_NORMAL)
    pb.stage("depth_normal")
    push_state(ctx, pb, ("planes",))
    ctx.problem_stage(STAGE.DEPTH_TO_WEAK)
    pb.stage("depth_to_weak", None)
    st = pull_state(ctx)
    same = (st["weak_info"] == pb.weak_info).mean()
    print("DepthToWeak with labels: states identical %.5f  hist gpu %s oracle %s" % (
        same, np.bincount(st["weak_info"].ravel(), minlength=3), np.bincount(pb.weak_info.ravel(), minlength=3)))
    assert same >= 0.9995  # measured 0.99996
    push_state(ctx, pb, ("weak_info",))
    ctx.problem_stage(STAGE.LOCAL_REFINE)
    pb.stage("local_refine")
    st = pull_state(ctx)
    with np.errstate(all="ignore"):
        rel = np.abs(st["planes"][..., 3] - pb.planes[..., 3]) / np.abs(pb.planes[..., 3])
    print("LocalRefine with labels: depth within 1e-4: %.5f" % (rel <= 1e-4).mean())
    assert (rel <= 1e-4).mean() >= 0.9995  # measured 1.0
    ctx.problem_finish()
    ctx.view_set_sa_mask(1, None)


def test_pass_with_labels_against_reference_binary(ctx, office_sa):
    """one whole use_APD pass with a label map: the product against the REFERENCE's own RunPatchMatch on the same inputs
    (statistical: different RNGs), beside the reference against itself with another seed"""
    from oracle import ref_binding as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libapd_ref.so not built (needs /root/reference at build time)")
    from apde_mvs_b200.binding import FIELD
    scene, sched = office_sa
    v = 3
    lab = make_labels(256, 192, 21)
    _setup(ctx, sched, v, lab)
    cams, prm = ctx.problem_cameras()
    _, _, n = ctx.problem_dims()
    imgs = [ctx.problem_image(i) for i in range(n)]
    depths = [d for d in ctx.problem_get(FIELD.SRC_DEPTH)]
    st = pull_state(ctx)
    # planes at entry = (world normal, depth) at the working size, with the weak / confidence maps (APD.cpp:614-683)
    planes_in, w0, c0 = st["planes"].copy(), st["weak_info"].copy(), st["confidence"].copy()
    ctx.problem_run()
    ctx.problem_finish()
    depth, _, weak, _ = ctx.view_download(v)
    ref.set_sa_mask(lab)
    try:
        rp, rweak, _, _ = ref.run_pass(imgs, [cams[i] for i in range(n)], prm, planes=planes_in, weak=w0, conf=c0, depths=depths, seed=31337)
        rp2, rweak2, _, _ = ref.run_pass(imgs, [cams[i] for i in range(n)], prm, planes=planes_in, weak=w0, conf=c0, depths=depths, seed=777)
    finally:
        ref.set_sa_mask(None)
    rp_plain, rweak_plain, _, _ = ref.run_pass(imgs, [cams[i] for i in range(n)], prm, planes=planes_in, weak=w0, conf=c0, depths=depths, seed=31337)
    rdepth = rp[..., 3]
    gt = scene.gt_depth[v]
    inner = np.zeros_like(gt, bool)
    inner[12:-12, 12:-12] = True
    sel = inner & (gt > 0) & (weak == 1) & (rweak == 1)
    sel2 = inner & (gt > 0) & (rweak2 == 1) & (rweak == 1)
    with np.errstate(all="ignore"):
        both = (np.abs(depth - rdepth) <= 0.01 * rdepth)[sel].mean()
        self_agree = (np.abs(rp2[..., 3] - rdepth) <= 0.01 * rdepth)[sel2].mean()
        acc_g = (np.abs(depth - gt) <= 0.01 * gt)[sel].mean()
        acc_r = (np.abs(rdepth - gt) <= 0.01 * gt)[sel].mean()
    hg = np.bincount(weak[inner], minlength=3) / inner.sum()
    hr = np.bincount(rweak[inner], minlength=3) / inner.sum()
    hp = np.bincount(rweak_plain[inner], minlength=3) / inner.sum()
    print("pass with labels: %d common strong px; within 1%% of the reference %.4f (reference vs itself %.4f); GT accuracy ours %.4f "
          "reference %.4f; state shares ours %s reference %s (reference without the map %s)" % (
              sel.sum(), both, self_agree, acc_g, acc_r, np.round(hg, 3), np.round(hr, 3), np.round(hp, 3)))
    assert both >= self_agree - 0.015
    assert acc_g >= acc_r - 0.015
    assert np.abs(hg - hr).max() < 0.03
    ctx.view_set_sa_mask(v, None)
