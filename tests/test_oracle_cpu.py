"""CPU tests (no GPU): the oracle against golden vectors produced by the REFERENCE's own device functions
(tests/golden/ref_costs.npz, generated on a B200 by tests/golden/make_ref_golden.py), RNG known answers, scheduling
invariants, the C-ABI export table and the multi-GPU host logic (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, orc, oracle_problem, ref_params

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_costs.npz")


def _golden_problem(z, use_apd=0):
    cams = []
    for row in z["cams"]:
        cam = orc.OCamera()
        C.memmove(C.byref(cam), row.tobytes(), 120)
        cams.append(cam)
    p = ref_params(use_apd=use_apd)
    p.depth_min, p.depth_max = float(z["iparams"][0]), float(z["iparams"][1])
    return orc.Problem(list(z["images"]), cams, p, depths=list(z["depths"]), seed=7, stream=2, tex_mode=1), cams


def _stats(name, got, want):
    d = np.abs(got - want)
    print("%s: n=%d max %.3g p99 %.3g frac<=1e-4 %.5f frac<=1e-3 %.5f" % (name, d.size, d.max(), np.quantile(d, 0.99), (d <= 1e-4).mean(), (d <= 1e-3).mean()))
    return d


def test_oracle_matches_reference_ncc_old():
    """ComputeBilateralNCCOld (APD.cu:596-663) as executed by the reference binary on a B200"""
    z = np.load(GOLDEN)
    pb, _ = _golden_problem(z)
    got = pb.eval_costs(z["old_tuples"], z["old_planes"], 0)
    d = _stats("NCC-Old oracle vs reference", got, z["old_costs"])
    # measured: 99.90 % within 1e-4, max 5.4e-4, p99 1e-5.  The oracle follows the reference build's operation order; what is
    # left is MUFU.RCP (<= 1 ulp off the correctly rounded reciprocal the oracle uses) moving a sample across a 1/256 weight
    # bucket.  The product, which has the same MUFU unit, matches these vectors bit for bit (tests/test_gpu_reference_pins.py).
    assert (d <= 1e-4).mean() >= 0.998 and (d <= 1e-3).all()
    assert ((got == 2.0) == (z["old_costs"] == 2.0)).all()
    # the texture unit's weight quantisation is part of the reference's formulation: exact bilinear must fit worse
    pb.pb.tex_mode = 0
    d0 = np.abs(pb.eval_costs(z["old_tuples"], z["old_planes"], 0) - z["old_costs"])
    assert np.quantile(d0, 0.9) > np.quantile(d, 0.9)


def test_oracle_matches_reference_geom_cost():
    """ComputeGeomConsistencyCost (APD.cu:865-902)"""
    z = np.load(GOLDEN)
    pb, _ = _golden_problem(z)
    got = pb.eval_costs(z["old_tuples"], z["old_planes"], 2)
    d = _stats("geom oracle vs reference", got, z["geom_costs"])
    assert (d <= 1e-4).all()  # measured: max 4.1e-5 (MUFU.RCP / MUFU.SQRT vs correctly rounded)
    assert ((got == 3.0) == (z["geom_costs"] == 3.0)).all()


def test_oracle_matches_reference_ncc_new():
    """ComputeBilateralNCCNew (APD.cu:448-593): centre patch + focal-weighted anchor patches"""
    z = np.load(GOLDEN)
    pb, _ = _golden_problem(z, use_apd=1)
    pb.weak_info[...] = z["new_weak"]
    pb.anchors[...] = z["new_anchors"]
    pb.selected_views[...] = z["new_sel"]
    got = pb.eval_costs(z["new_tuples"], z["new_planes"], 1)
    d = _stats("NCC-New oracle vs reference", got, z["new_costs"])
    assert (d <= 1e-4).mean() >= 0.995 and (d <= 1e-3).all()  # measured: 99.65 % within 1e-4, max 5.0e-4


GOLDEN_SA = os.path.join(ROOT, "tests", "golden", "ref_costs_sa.npz")


def test_oracle_matches_reference_with_segment_labels():
    """NCC-Old branch B (APD.cu:664-719) and the label tests of NCC-New (APD.cu:493-497, 526-530) as executed by the reference
    binary on a B200 with a label map in sa_mask_cuda (tests/golden/make_ref_golden_sa.py)"""
    z, zs = np.load(GOLDEN), np.load(GOLDEN_SA)
    pb, _ = _golden_problem(z)
    pb.set_sa_mask(zs["labels"])
    got = pb.eval_costs(z["old_tuples"], z["old_planes"], 0)
    d = _stats("NCC-Old + labels oracle vs reference", got, zs["old_costs"])
    assert (d <= 1e-4).mean() >= 0.997 and (d <= 1e-3).all()  # measured: 99.87 % within 1e-4, max 8.8e-4
    assert (np.abs(zs["old_costs"] - z["old_costs"]) > 1e-3).mean() > 0.2  # the map matters on this input
    pb, _ = _golden_problem(z, use_apd=1)
    pb.set_sa_mask(zs["labels"])
    pb.weak_info[...] = z["new_weak"]
    pb.anchors[...] = z["new_anchors"]
    pb.selected_views[...] = z["new_sel"]
    got = pb.eval_costs(z["new_tuples"], z["new_planes"], 1)
    d = _stats("NCC-New + labels oracle vs reference", got, zs["new_costs"])
    assert (d <= 1e-4).mean() >= 0.996 and (d <= 1e-3).all()  # measured: 99.77 % within 1e-4, max 3.2e-4
    assert (np.abs(zs["new_costs"] - z["new_costs"]) > 1e-3).mean() > 0.05


def test_oracle_segment_label_properties():
    """size-independent properties of the label path: an all-zero map is no map; one segment covering the image leaves NCC-New
    unchanged (same taps, same order) and turns NCC-Old into the quadrant walk over the same 36 taps (same samples, another
    summation order: equal up to fp32 rounding on textured patches)"""
    z = np.load(GOLDEN)
    pb, _ = _golden_problem(z)
    h, w = z["images"][0].shape
    base = pb.eval_costs(z["old_tuples"], z["old_planes"], 0)
    pb.set_sa_mask(np.zeros((h, w), np.uint8))
    assert np.array_equal(pb.eval_costs(z["old_tuples"], z["old_planes"], 0), base)
    pb.set_sa_mask(np.full((h, w), 7, np.uint8))
    one = pb.eval_costs(z["old_tuples"], z["old_planes"], 0)
    t = z["old_tuples"]
    inner = (t[:, 0] >= 5) & (t[:, 0] < w - 5) & (t[:, 1] >= 5) & (t[:, 1] < h - 5)  # branch B skips taps outside the image
    ok = inner & (base < 1.9)
    assert np.abs(one - base)[ok].max() < 5e-3 and (np.abs(one - base)[ok] <= 1e-4).mean() > 0.9
    pb, _ = _golden_problem(z, use_apd=1)
    pb.weak_info[...] = z["new_weak"]
    pb.anchors[...] = z["new_anchors"]
    pb.selected_views[...] = z["new_sel"]
    base = pb.eval_costs(z["new_tuples"], z["new_planes"], 1)
    pb.set_sa_mask(np.full((h, w), 7, np.uint8))
    assert np.array_equal(pb.eval_costs(z["new_tuples"], z["new_planes"], 1), base)


def test_oracle_pass_vs_reference_pass_statistical():
    """whole photometric pass: the reference (XORWOW, seed patched) and the oracle (Philox) are different random
    processes, so parity is statistical: on textured pixels both must sit within 1 % of the ground truth"""
    z = np.load(GOLDEN)
    pb, cams = _golden_problem(z)
    pb.pb.num_threads = 8
    pb.stage("run_pass")
    gt = z["gt_depth"]
    ref_depth, orc_depth = z["pass_planes"][..., 3], pb.planes[..., 3]
    m = 8
    inner = np.zeros_like(gt, bool)
    inner[m:-m, m:-m] = True
    sel = inner & (gt > 0) & (z["pass_weak"] == 1) & (pb.weak_info == 1)
    ref_ok = np.abs(ref_depth - gt)[sel] <= 0.01 * gt[sel]
    orc_ok = np.abs(orc_depth - gt)[sel] <= 0.01 * gt[sel]
    both = np.abs(orc_depth - ref_depth)[sel] <= 0.01 * ref_depth[sel]
    print("strong pixels %d: reference within 1%% of GT %.4f, oracle %.4f, oracle within 1%% of reference %.4f" % (
        sel.sum(), ref_ok.mean(), orc_ok.mean(), both.mean()))
    assert sel.sum() > 2000
    assert orc_ok.mean() >= ref_ok.mean() - 0.02
    assert both.mean() >= 0.97
    # classification statistics agree
    h_ref = np.bincount(z["pass_weak"][inner], minlength=3) / inner.sum()
    h_orc = np.bincount(pb.weak_info[inner], minlength=3) / inner.sum()
    print("state shares ref %s oracle %s" % (np.round(h_ref, 3), np.round(h_orc, 3)))
    assert np.abs(h_ref - h_orc).max() < 0.05


# ------------------------------------------------------------------------------------------------ RNG
def _philox_py(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10, and an independent pure-Python restatement"""
    assert _philox_py([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _philox_py([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    rng = np.random.default_rng(0)
    for _ in range(50):
        seed, stream, pixel, site, block = [int(x) for x in rng.integers(0, 2 ** 32, 5, dtype=np.uint64)]
        got = [int(x) for x in orc.philox(seed, stream, pixel, site, block)]
        assert got == _philox_py([pixel, site, block, 0], [seed, stream])


# ------------------------------------------------------------------------------------------------ scheduling / indexing
def test_checkerboard_candidates_are_opposite_colour_and_in_bounds():
    rng = np.random.default_rng(1)
    for (w, h) in ((37, 29), (64, 48)):
        costs = rng.random((h, w)).astype(np.float32)
        for _ in range(300):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            pos, flags = orc.checkerboard_candidates(costs, x, y)
            expect = [y > 0, y > 2, y < h - 1, y < h - 3, x > 0, x > 2, x < w - 1, x < w - 3]
            assert list(flags.astype(bool)) == expect
            for k in range(8):
                if flags[k]:
                    py, px = divmod(int(pos[k]), w)
                    assert 0 <= px < w and 0 <= py < h
                    assert ((px + py) & 1) != ((x + y) & 1)  # red/black invariant: in-place update is race free
    # far arm picks the minimum cost among its 11 taps; ties keep the nearest (strict <)
    costs = np.ones((64, 64), np.float32)
    costs[20 - 3 - 2 * 4, 30] = 0.5
    pos, _ = orc.checkerboard_candidates(costs, 30, 20)
    assert pos[1] == (20 - 11) * 64 + 30
    costs[:] = 1.0
    pos, _ = orc.checkerboard_candidates(costs, 30, 20)
    assert pos[1] == (20 - 3) * 64 + 30


def test_quirk_invalid_neighbour_blocks_propagation():
    """quirk 2: zero-initialised cost rows of invalid neighbours win FindMinCostIndex, so a border
    pixel never takes a propagated plane (only refinement).  Pixel (0, y): left neighbours are invalid."""
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(96, 72, num_views=3, num_src=2, seed=2)
    pb = oracle_problem(scene, 1, ref_params(), threads=4)
    pb.stage("random_init")
    before = pb.planes.copy()
    pb.stage("propagate_strong", 0, 0)
    pb.stage("propagate_strong", 0, 1)
    # interior pixels overwhelmingly adopt a neighbour's plane at iteration 0; border-column pixels can only change through
    # the five refinement candidates, which perturb their own plane or draw a random one -- never copy a neighbour
    nb_planes = {tuple(np.round(before[y, x], 6)) for y in range(72) for x in range(0, 30)}
    copied_border = sum(tuple(np.round(pb.planes[y, 0], 6)) in nb_planes and not np.allclose(pb.planes[y, 0], before[y, 0]) for y in range(4, 68))
    copied_inner = sum(tuple(np.round(pb.planes[y, 8], 6)) in nb_planes and not np.allclose(pb.planes[y, 8], before[y, 8]) for y in range(4, 68))
    print("pixels that copied a neighbour plane: border column %d, interior column %d" % (copied_border, copied_inner))
    assert copied_border == 0 and copied_inner >= 5


def test_half_grid_row_limit_quirk():
    """quirk 7: for odd H with (H/2) % 16 == 0 the last row is never visited by the red/black kernels"""
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(48, 33, num_views=3, num_src=2, seed=3)
    pb = oracle_problem(scene, 1, ref_params(), threads=4)
    pb.stage("random_init")
    before = pb.planes.copy()
    for c in (0, 1):
        pb.stage("propagate_strong", 0, c)
    assert np.array_equal(pb.planes[32], before[32])
    assert not np.array_equal(pb.planes[31], before[31])


# ------------------------------------------------------------------------------------------------ fusion oracle
def test_fusion_oracle_on_ground_truth_maps():
    """perfect depth maps fuse into points on the surface; every accepted point masks the pixels it consumed"""
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(96, 72, num_views=4, num_src=3, seed=5, with_color=True)
    V = 4
    depths = np.stack(scene.gt_depth).astype(np.float32)
    n_w = np.array([-0.15, 0.1, 1.0]); n_w /= -np.linalg.norm(n_w)
    normals = np.tile(n_w.astype(np.float32), (V, 72, 96, 1))
    weaks = np.ones((V, 72, 96), np.uint8)
    confs = np.ones((V, 72, 96), np.uint8)
    xyz, bgr, skip = orc.fusion(scene.cameras, depths, normals, weaks, confs, scene.pairs, np.stack(scene.colors))
    assert skip.sum() == 0  # no WEAK pixels -> nothing to filter
    assert 0.2 * 96 * 72 < len(xyz) < V * 96 * 72
    plane_res = np.abs(4 + 0.15 * xyz[:, 0] - 0.1 * xyz[:, 1] - xyz[:, 2])
    assert np.quantile(plane_res, 0.99) < 1e-2
    assert bgr.min() >= 0 and bgr.max() <= 255
    # empty input: all depths invalid -> no points
    xyz0, _, _ = orc.fusion(scene.cameras, np.zeros_like(depths), normals, weaks, confs, scene.pairs)
    assert len(xyz0) == 0


def test_fusion_oracle_tat_variants():
    """RunFusion_TAT_I / _A (APD.cpp:1229-1608): k-of-N consistency with thresholds growing in k; diff[] carries over
    between pixels of a view (quirk 14), so a pixel whose neighbours all fall outside can still be accepted"""
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(96, 72, num_views=4, num_src=3, seed=5, with_color=True)
    V = 4
    depths = np.stack(scene.gt_depth).astype(np.float32)
    n_w = np.array([-0.15, 0.1, 1.0]); n_w /= -np.linalg.norm(n_w)
    normals = np.tile(n_w.astype(np.float32), (V, 72, 96, 1))
    weaks = np.ones((V, 72, 96), np.uint8)
    confs = np.ones((V, 72, 96), np.uint8)
    cols = np.stack(scene.colors)
    counts = {}
    for variant in (1, 2):
        xyz, bgr, _ = orc.fusion(scene.cameras, depths, normals, weaks, confs, scene.pairs, cols, variant=variant)
        counts[variant] = len(xyz)
        assert len(xyz) > 0.5 * 96 * 72
        res = np.abs(4 + 0.15 * xyz[:, 0] - 0.1 * xyz[:, 1] - xyz[:, 2])
        assert np.quantile(res, 0.99) < 1e-2
        assert bgr.min() >= 0 and bgr.max() <= 255
    # TAT_A has the looser depth threshold and no angle test: it accepts at least as many points on clean maps
    assert counts[2] >= counts[1]
    # carry-over: blank the neighbours' depth everywhere except one early row -> later pixels of view 0 reuse the
    # measurements of that row and are still accepted (a per-pixel diff[] would reject them all)
    d2 = depths.copy()
    d2[1:, 8:, :] = 0
    xyz, _, _ = orc.fusion(scene.cameras, d2, normals, weaks, confs, scene.pairs, cols, variant=2)
    assert len(xyz) > 96 * 20, "stale diff[] must keep accepting pixels of view 0 (got %d)" % len(xyz)
    # a view with fewer than two neighbours never fuses anything (k starts at 2)
    xyz1, _, _ = orc.fusion(scene.cameras, depths, normals, weaks, confs, [p[:1] for p in scene.pairs], cols, variant=1)
    assert len(xyz1) == 0


# ------------------------------------------------------------------------------------------------ C ABI
def test_c_abi_exports_every_declared_symbol(apde_lib):
    hdr = open(os.path.join(ROOT, "include", "apde.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(apde_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    from apde_mvs_b200.binding import lib_path
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path()], text=True)
    exported = set(re.findall(r"\sT\s+(apde_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/apde.h but not exported: %s" % missing
    # struct layouts the reference interface fixes
    from apde_mvs_b200.binding import Camera, Params
    assert C.sizeof(Camera) == 120  # main.h:50-61
    assert C.sizeof(Params) == 72


def test_header_is_plain_c_and_struct_sizes_match_the_binding(apde_lib, tmp_path):
    """include/apde.h is the drop-in boundary: it must compile as C (no C++ types across it), link against libapde.so from a C
    program, and the ctypes mirrors in apde_mvs_b200/binding.py must have the compiler's struct sizes"""
    from apde_mvs_b200.binding import Camera, Params, Schedule, Timing, lib_path
    src = tmp_path / "abi.c"
    src.write_text('''#include <stdio.h>
#include "apde.h"
int main(void) {
    apde_params p;
    apde_schedule s;
    apde_params_default(&p);
    apde_schedule_default(&s);
    printf("%zu %zu %zu %zu %d %d %d %s\\n", sizeof(apde_camera), sizeof(apde_params), sizeof(apde_schedule), sizeof(apde_timing),
           p.top_k, s.geom_iterations, s.use_sa, apde_version());
    return apde_view_set_sa_mask(0, 0, 0, 0, 0) == 0;  /* null context: an error code, not a crash */
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + libdir, "-lapde", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    f = out.stdout.split()
    assert [int(x) for x in f[:4]] == [C.sizeof(Camera), C.sizeof(Params), C.sizeof(Schedule), C.sizeof(Timing)]
    assert f[4:7] == ["4", "3", "1"] and "sm_100a" in out.stdout


def test_no_cpu_fallback(apde_lib):
    """without a CUDA device the product fails loudly instead of computing on the host"""
    from apde_mvs_b200.binding import ApdeError, Context
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(ApdeError, match="no CPU fallback"):
        Context(0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "apde_mvs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/apd_oracle.cpp", "").replace("Same spec as", "") or f == "scene.py", f


# ------------------------------------------------------------------------------------------------ multi-GPU host logic
def test_view_sharding_partition():
    from apde_mvs_b200.sharding import shard
    for V in (1, 7, 11, 26, 300):
        for world in (1, 2, 4, 8):
            blocks = [shard(V, world, r) for r in range(world)]
            covered = [v for f, c in blocks for v in range(f, f + c)]
            assert covered == list(range(V))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _gloo_job_worker(rank, world, port, V, q):
    """host logic of a multi-GPU job on CPU (gloo, world_size 2): the rendezvous hands the SAME 128-byte id to every rank, every
    rank joins with its own rank / world, the schedule and the fusion are issued as collective calls in the same order"""
    import torch.distributed as dist
    from apde_mvs_b200.binding import COMM_ID_BYTES, Schedule
    from apde_mvs_b200.distributed import Job
    from apde_mvs_b200.sharding import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeCtx:  # stands in for binding.Context: records what the launcher-side logic asks of the library
        made = 0

        def __init__(self):
            self.log = []

        @staticmethod
        def comm_create_id():
            FakeCtx.made += 1
            return bytes((7 * i + 3) % 256 for i in range(COMM_ID_BYTES))

        def comm_init(self, comm_id, r, w):
            self.log.append(("init", bytes(comm_id), r, w))

        def comm_info(self):
            f, c = shard(V, world, rank)
            return rank, world, f, c

        def num_passes(self, sched):
            return 3

        def run_schedule_pass(self, sched, p, t):
            assert sched.num_views_local == 0  # the library deals the views out itself
            self.log.append("pass%d" % p)

        def fuse_collective(self, weak_filter, variant=0):
            self.log.append("fuse%d%d" % (int(weak_filter), variant))
            return ("xyz", "bgr") if rank == 0 else (None, None)

    ok = True
    try:
        ctx = FakeCtx()
        job = Job(ctx, dist)
        job.run_schedule(Schedule())
        xyz, _ = job.fuse(True, variant=1)
        want_id = bytes((7 * i + 3) % 256 for i in range(COMM_ID_BYTES))
        ok = ctx.log[0] == ("init", want_id, rank, world) and ctx.log[1:] == ["pass0", "pass1", "pass2", "fuse11"]
        ok = ok and FakeCtx.made == (1 if rank == 0 else 0)  # only rank 0 creates the id
        ok = ok and (job.first, job.count) == shard(V, world, rank)
        ok = ok and ((xyz == "xyz") if rank == 0 else (xyz is None))
    except AssertionError as e:  # report instead of hanging the peer
        ok = False
        print("rank %d: %r" % (rank, e))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("V", [6, 5])
def test_multi_gpu_job_host_logic_gloo_world2(V):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + V
    procs = [ctx.Process(target=_gloo_job_worker, args=(r, 2, port, V, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_view_blocks_of_the_library_match_the_host_mirror(apde_lib):
    """apde_comm_block_of (what the NCCL exchange inside libapde uses) == sharding.shard, and the blocks tile [0, V)"""
    from apde_mvs_b200.sharding import shard
    for V in (0, 1, 5, 7, 26, 88, 300):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for rank in range(world):
                f, c = C.c_int(), C.c_int()
                assert apde_lib.apde_comm_block_of(V, world, rank, C.byref(f), C.byref(c)) == 0
                assert (f.value, c.value) == shard(V, world, rank)
                assert f.value == nxt
                nxt += c.value
            assert nxt == V
    f, c = C.c_int(), C.c_int()
    assert apde_lib.apde_comm_block_of(4, 2, 2, C.byref(f), C.byref(c)) != 0  # rank out of range
