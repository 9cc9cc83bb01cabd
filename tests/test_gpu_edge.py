"""GPU edge cases: error behaviour of the C ABI, odd sizes (quirk 7), one source view, many source views."""
import numpy as np
import pytest

from helpers import oracle_from_ctx, plane_depth, pull_state, ref_params, to_apde_params

pytestmark = pytest.mark.gpu


def test_error_behaviour(apde_lib):
    """bad calls return an error code + message (the reference prints and exit()s, APD.cpp:417-450, 528-531)"""
    from apde_mvs_b200.binding import ApdeError, Context, default_params
    from apde_mvs_b200.scene import make_plane_scene
    c = Context(0)
    p = default_params()
    p.use_APD, p.state = 0, 0
    with pytest.raises(ApdeError, match="not committed"):
        c.problem_setup(0, p, 1, 1)
    scene = make_plane_scene(96, 64, num_views=3, num_src=2, seed=1)
    c.scene_begin(3, 96, 64)
    for v in range(3):
        c.scene_set_view(v, scene.images[v], scene.cameras[v])
        c.scene_set_pairs(v, scene.pairs[v])
    with pytest.raises(ApdeError, match="so much images"):
        c.scene_set_pairs(0, list(range(32)))
    with pytest.raises(ApdeError, match="out of range"):
        c.scene_set_pairs(0, [7])
    c.scene_commit()
    with pytest.raises(ApdeError, match="bad reference view"):
        c.problem_setup(9, p, 1, 1)
    p2 = default_params()
    p2.use_APD, p2.state = 1, 1
    with pytest.raises(ApdeError, match="no depth map|weak/confidence|previous depth"):
        c.problem_setup(0, p2, 1, 1)  # round >= 1 needs the previous round's maps
    p3 = default_params()
    p3.use_APD, p3.state, p3.strong_increment = 0, 0, 3
    with pytest.raises(ApdeError, match="patch geometry"):
        c.problem_setup(0, p3, 1, 1)
    c.problem_setup(0, p, 1, 1)
    with pytest.raises(ApdeError, match="out of range"):
        c.eval_costs(np.array([[200, 3, 1]]), np.array([[0, 0, -1, 4.0]], np.float32), 0)
    with pytest.raises(ApdeError):  # fusion without a scene
        Context(0).fuse()
    c.close()


@pytest.mark.parametrize("size,views,src", [((131, 97), 4, 3), ((48, 33), 3, 2), ((96, 64), 3, 1), ((80, 64), 15, 13),
                                            ((64, 48), 32, 31)])
def test_whole_pass_matches_oracle_on_odd_shapes(ctx, size, views, src):
    """odd widths / heights (incl. H = 33: the last row is never visited by the red/black kernels, quirk 7), a single
    source view, N = 13 (the shared-memory sizing switches to 64-thread blocks) and N = 31, the MAX_IMAGES limit (main.h:40)"""
    from apde_mvs_b200.scene import make_plane_scene
    w, h = size
    scene = make_plane_scene(w, h, num_views=views, num_src=src, seed=5)
    ctx.load_scene(scene)
    p = to_apde_params(ref_params())
    ctx.problem_setup(1, p, 1, 4321)
    pb = oracle_from_ctx(ctx, 4321, 1, threads=8)
    cams, _ = ctx.problem_cameras()
    ctx.problem_run()
    pb.stage("run_pass")
    st = pull_state(ctx)
    dg, do = st["planes"][..., 3], pb.planes[..., 3]
    with np.errstate(all="ignore"):
        ok = (np.abs(dg - do) <= 0.01 * np.abs(do)) | (np.isnan(dg) & np.isnan(do))
    m = 8
    frac = ok[m:-m, m:-m].mean()
    print("%s N=%d: depth within 1%% of the oracle pass %.4f; states equal %.4f" % (size, src, frac, (st["weak_info"] == pb.weak_info).mean()))
    assert frac >= 0.97
    assert (st["weak_info"] == pb.weak_info).mean() >= 0.97
    if h == 33:  # rows >= 32 are outside the half grid: only init + the full-grid tail stages touch them
        assert ok[32].mean() >= 0.9
    ctx.problem_finish()
    d, n, wk, cf = ctx.view_download(1)
    assert d.shape == (h, w) and n.shape == (h, w, 3)


def test_fusion_variants_on_uploaded_maps_with_holes():
    """all three fusion variants on uploaded maps with holes, noise and mixed WEAK/STRONG labels == the oracle, point for
    point and in the reference's order (APD.cpp:1051-1608); holes exercise the diff[] carry-over of the TaT variants"""
    from apde_mvs_b200.binding import Context
    from apde_mvs_b200.scene import make_plane_scene
    from oracle import binding as orc
    W, H, V = 160, 120, 5
    scene = make_plane_scene(W, H, num_views=V, num_src=4, seed=9, with_color=True)
    rng = np.random.default_rng(3)
    depths = np.stack(scene.gt_depth).astype(np.float32)
    depths *= (1 + rng.normal(0, 2e-4, depths.shape)).astype(np.float32)  # noise around the TaT depth thresholds (1/3500)
    holes = rng.random(depths.shape) < 0.15
    depths[holes] = 0
    depths[:, 40:60, 30:90] = 0  # a block hole: long carry-over runs
    n_w = np.array([-0.15, 0.1, 1.0]); n_w /= -np.linalg.norm(n_w)
    normals = np.tile(n_w.astype(np.float32), (V, H, W, 1))
    normals += rng.normal(0, 0.03, normals.shape).astype(np.float32)
    weaks = (rng.random((V, H, W)) < 0.7).astype(np.uint8)
    confs = rng.integers(0, 255, (V, H, W)).astype(np.uint8)
    ctx = Context(0)
    ctx.load_scene(scene)
    for v in range(V):
        ctx.view_upload(v, depths[v], normals[v], weaks[v], confs[v])
    cols = np.stack(scene.colors)
    for variant in (0, 1, 2):
        for wf in (True, False):
            xyz_o, bgr_o, _ = orc.fusion(scene.cameras, depths, normals, weaks, confs, scene.pairs, cols, weak_filter=wf,
                                         variant=variant)
            xyz_g, bgr_g = ctx.fuse(wf, variant=variant)
            print("variant %d weak_filter %d: gpu %d oracle %d points" % (variant, wf, len(xyz_g), len(xyz_o)))
            assert len(xyz_o) > 500
            assert len(xyz_g) == len(xyz_o)
            assert np.allclose(xyz_g, xyz_o, rtol=1e-5, atol=1e-5)
            assert np.abs(bgr_g - bgr_o).max() <= 1e-3
    # the cloud of a count-only call is handed over once (apde_fuse_take_points); the direct call with buffers gives the same
    # points, truncates at max_points, and leaves nothing to take
    from apde_mvs_b200.binding import ApdeError
    xyz_a, bgr_a = ctx.fuse(True, variant=0)
    with pytest.raises(ApdeError):
        ctx._take_points(len(xyz_a))
    xyz_b, bgr_b = ctx.fuse(True, max_points=len(xyz_a), variant=0)
    assert np.array_equal(xyz_a, xyz_b) and np.array_equal(bgr_a, bgr_b)
    xyz_c, _ = ctx.fuse(True, max_points=100, variant=0)
    assert len(xyz_c) == 100 and np.array_equal(xyz_c, xyz_a[:100])
    with pytest.raises(ApdeError):
        ctx._take_points(10)
    ctx.close()


def test_banded_sweep_is_identical():
    """DepthToWeak / LocalRefine column storage processed in bands of rows (large images, apde_set_sweep_budget_mb) gives
    bit-identical maps to the single-band path, in photometric, geometric and use_APD passes"""
    from apde_mvs_b200.binding import Context, default_schedule
    from apde_mvs_b200.scene import make_office_scene
    scene = make_office_scene(320, 240, num_views=4, num_src=3, seed=3, weak=0.3)
    sched = default_schedule()
    sched.rounds, sched.geom_iterations, sched.seed = 2, 1, 9
    out = []
    for mb in (0, 1):  # 1 MB: 320 x 3 views x 62 slots x 4 B (x 2 geometric) = 238 / 476 KB per row -> 4 / 2 rows per band
        ctx = Context(0)
        ctx.set_sweep_budget_mb(mb)
        ctx.load_scene(scene)
        t = ctx.run_schedule(sched)
        out.append(([ctx.view_download(v) for v in range(4)], t))
        ctx.close()
    (a, ta), (b, tb) = out
    assert tb.kernel_launches > ta.kernel_launches  # the banded run really took the banded path
    assert (ta.evals_ncc_old, ta.evals_ncc_new) != (0, 0)
    for v in range(4):
        for k in range(4):
            assert np.array_equal(a[v][k], b[v][k]), "view %d map %d differs between banded and single-band sweeps" % (v, k)


def test_strong_propagation_kernel_variants_identical(tmp_path):
    """the warp-compacted refinement of k_prop_strong (pairs of (pixel, selected view) dealt out view-major across the
    warp) and the per-lane twin k_prop_strong_v1 give bit-identical maps over a whole two-round schedule"""
    import os
    import subprocess
    import sys
    from helpers import ROOT
    outs = []
    for v1 in ("0", "1"):
        out = str(tmp_path / ("maps_v1_%s.npz" % v1))
        env = dict(os.environ, APDE_STRONG_V1=v1)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_maps.py"), out, "320", "240", "5", "4", "0.25", "2"], env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        outs.append(np.load(out))
    a, b = outs
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 20
    for k in a.files:
        assert np.array_equal(a[k], b[k]), "map %s differs between the compacted and the per-lane kernel" % k


def test_segment_label_kernel_variants_identical(tmp_path):
    """with segment-label maps on every view, the default kernels (SA twins of the column pipeline and of the sweep columns)
    and the thread-per-pixel SA twins (APDE_LEGACY_PROP / APDE_LEGACY_SWEEP) give bit-identical maps over a two-round
    schedule -- and the labels change the result"""
    import os
    import subprocess
    import sys
    from helpers import ROOT
    outs = []
    for env_extra in ({"APDE_DUMP_SA": "1"}, {"APDE_DUMP_SA": "1", "APDE_LEGACY_PROP": "1", "APDE_LEGACY_SWEEP": "1"}, {}):
        out = str(tmp_path / ("maps_%d.npz" % len(outs)))
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_maps.py"), out, "320", "240", "5", "4", "0.25", "2"],
                              env=dict(os.environ, **env_extra), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        outs.append(np.load(out))
    a, b, plain = outs
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 20
    for k in a.files:
        assert np.array_equal(a[k], b[k]), "map %s differs between the default and the thread-per-pixel kernels under labels" % k
    assert any(not np.array_equal(a[k], plain[k]) for k in a.files if k.startswith("w"))


def test_weak_propagation_pipelines_identical(tmp_path):
    """weak pixels: the anchor-sorted pipeline (default: anchor patches gathered by sorted (pixel, anchor) pairs, centre patches
    and the softmax mix in separate kernels), the (pixel, view) column kernels (APDE_WEAK_COLUMNS=1) and the thread-per-pixel
    kernel (APDE_LEGACY_PROP=1) give bit-identical maps over a two-round schedule with 30 % weak texture"""
    import os
    import subprocess
    import sys
    from helpers import ROOT
    outs = []
    for env_extra in ({}, {"APDE_WEAK_COLUMNS": "1"}, {"APDE_LEGACY_PROP": "1"}):
        out = str(tmp_path / ("maps_%d.npz" % len(outs)))
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_maps.py"), out, "400", "300", "5", "4", "0.3", "2"],
                              env=dict(os.environ, **env_extra), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        outs.append(np.load(out))
    a, b, c = outs
    assert len(a.files) == 20 and (a["w0"] == 0).mean() > 0.05  # there are WEAK pixels
    for k in a.files:
        assert np.array_equal(a[k], b[k]), "map %s differs between the anchor-sorted pipeline and the column kernels" % k
        assert np.array_equal(a[k], c[k]), "map %s differs between the anchor-sorted pipeline and the thread-per-pixel kernel" % k
