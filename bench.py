#!/usr/bin/env python
"""bench.py -- ref-views/s and cost evals/s of the APDe-MVS PatchMatch hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo (libapde.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own code (oracle/_ref) on the same scene

One "step" = the whole multi-scale schedule of main.cpp:303-367 (R rounds x (1 photometric + 3 geometric) passes) over
every reference view of a synthetic scene.  `value` = views completed per second of device time (inputs resident in HBM);
`e2e` = the same through the C ABI from pinned HOST buffers (scene upload + schedule + download of the four maps of every
view inside the timed region).

Workloads (--config):  headline = BASELINE.json's metric shape: 11 views per GPU at 1920x1080, 10 source views, 3 rounds;
                       C2 = ETH3D-office shape: 26 views 1550x1030, 10 source views, 2 rounds;
                       C3 = weak-texture shape: 6 views per GPU 1600x1200, 5 source views, 40 % weak-texture blobs.
N > 1 is ONE scene over N GPUs (a multi-GPU job of libapde, csrc/apde_comm.cu): the cameras stand on N arcs one above the
other, rank r owns arc r, and the nearest-camera source lists reach into the neighbouring arcs -- every geometric pass
reads depth maps that another rank produced in the previous pass and that NCCL carried over NVLink while the next view was
being computed.  --scaling weak (default): 11 views per GPU;  --scaling strong: a fixed scene of --views-total views.
torch.distributed (gloo) is used for the rendezvous only: it hands the 128-byte NCCL id around and reduces the timings.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_FLOP_PER_EVAL = 1340.0  # SURVEY.md 8(d) / BASELINE.md 2.4
ALG_L1_BYTES_PER_EVAL = 720.0
SAMPLES_PER_EVAL = 36.0
SAMPLES_PER_ANCHOR = 9.0  # NCC-New: 36 + 9 per valid anchor (APD.cu:521-533)
# dram__bytes_read.sum + dram__bytes_write.sum per full-resolution launch at the headline workload, from `ncu --set full`
# captures (profiles/): the dominant kernels are bound by the texture units, not by HBM.
NCU_TRAFFIC = {"prop_strong": (2.16e8, "k_prop_strong", "profiles/r02_ncu_top_kernels.md"),
               "depth_to_weak": (6.85e9, "k_sweep_columns", "profiles/r02_ncu_top_kernels.md")}
STAGE_NAMES = ["nearest_strong", "gen_anchors", "init", "prop_strong", "ransac_fit", "prop_weak", "depth_normal", "median",
               "depth_to_weak", "confidence", "local_refine"]
CONFIGS = {  # width, height, views per GPU, source views, rounds (0 = ComputeRoundNum), weak-texture share
    "headline": (1920, 1080, 11, 10, 0, 0.0),
    "C2": (1550, 1030, 26, 10, 0, 0.0),
    "C3": (1600, 1200, 6, 5, 0, 0.4),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--views-per-gpu", type=int, default=0)
    ap.add_argument("--src", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=-1, help="0 = reference rule (ComputeRoundNum)")
    ap.add_argument("--geom-iters", type=int, default=3)
    ap.add_argument("--weak", type=float, default=-1.0, help="share of every surface covered by weak-texture blobs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--views-total", type=int, default=88, help="--scaling strong: views of the fixed scene (a multiple of 11 arcs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fusion", action="store_true", help="skip the fusion report (outside the timed metric)")
    ap.add_argument("--no-job-check", action="store_true", help="N > 1: skip the bit-identity check against a single-GPU Jacobi run")
    ap.add_argument("--ref-views", type=int, default=2, help="reference arm: reference views timed per step (0 = all)")
    a = ap.parse_args()
    w, h, vpg, src, rounds, weak = CONFIGS[a.config]
    a.width, a.height = a.width or w, a.height or h
    a.views_per_gpu, a.src = a.views_per_gpu or vpg, a.src or src
    a.rounds = rounds if a.rounds < 0 else a.rounds
    a.weak = weak if a.weak < 0 else a.weak
    return a


# ----------------------------------------------------------------------------------------------- scene
def scene_layout(args, world):
    """(views per arc, arcs): weak scaling puts one arc of views_per_gpu views on every GPU; strong scaling fixes the scene"""
    if args.scaling == "strong":
        arcs = max(1, args.views_total // args.views_per_gpu)
        return args.views_per_gpu, arcs
    return args.views_per_gpu, world


def build_scene(args, world, rank=0, barrier=None):
    """the synthetic office scene; every rank renders its share of the views (a pool of processes) into /tmp, then all load all"""
    from apde_mvs_b200.scene import make_office_scene, office_setup, render_views
    per_arc, arcs = scene_layout(args, world)
    V = per_arc * arcs
    tag = "/tmp/apde_bench_%dx%d_%dx%d_%g" % (args.width, args.height, per_arc, arcs, args.weak)
    facets, K, poses = office_setup(args.width, args.height, per_arc, 2, args.weak, 60.0, arcs)
    mine = [v for v in range(V) if v % world == rank and not os.path.exists("%s_v%d.npz" % (tag, v))]
    if mine:
        for v, (gray, depth) in zip(mine, render_views(facets, K, poses, args.width, args.height, mine)):
            np.savez(tag + "_v%d.tmp.npz" % v, gray=gray, depth=depth)
            os.replace(tag + "_v%d.tmp.npz" % v, "%s_v%d.npz" % (tag, v))
    if barrier:
        barrier()
    pre = []
    for v in range(V):
        z = np.load("%s_v%d.npz" % (tag, v))
        pre.append((z["gray"], z["depth"]))
    return make_office_scene(args.width, args.height, per_arc, args.src, seed=2, weak=args.weak, prerendered=pre, rings=arcs)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "1000"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for nm, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_pinned = set()


def pin(arr):
    """cudaHostRegister a numpy buffer so the C ABI's host<->device copies use pinned memory"""
    if arr.ctypes.data in _pinned:
        return arr
    try:
        rt = C.CDLL("libcudart.so.12")
        rt.cudaHostRegister(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes), 0)
        rt.cudaGetLastError()  # a failed registration must not linger as the runtime's sticky last error
        _pinned.add(arr.ctypes.data)
    except Exception:
        pass
    return arr


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    from apde_mvs_b200.binding import Context, default_schedule, Timing
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")  # rendezvous only: the depth maps travel over NCCL inside libapde

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce_max(vals):
        if dist is None:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_sum(vals):
        if dist is None:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    scene = build_scene(args, world, rank, barrier)
    V = len(scene.images)
    images = [pin(np.ascontiguousarray(im)) for im in scene.images]
    ctx = Context(local_rank)
    if world > 1:
        from apde_mvs_b200.distributed import share_comm_id
        ctx.comm_init(share_comm_id(Context, dist), rank, world)

    def upload(c=ctx):
        c.scene_begin(V, scene.width, scene.height)
        for v in range(V):
            c.scene_set_view(v, images[v], scene.cameras[v])
            c.scene_set_pairs(v, scene.pairs[v])
        c.scene_commit()

    sched = default_schedule()
    sched.rounds, sched.geom_iterations, sched.seed = args.rounds, args.geom_iters, 1
    upload()
    _, _, first, count = ctx.comm_info()
    npass = ctx.num_passes(sched)
    cross = sum(1 for v in range(first, first + count) for s in scene.pairs[v] if not (first <= s < first + count))

    def step(timing, c=ctx, passes=None):
        for p in range(npass if passes is None else passes):
            c.run_schedule_pass(sched, p, timing)  # collective at N > 1: own block of views + depth-map exchange

    # ---- N > 1: the job must reproduce a single-GPU run with the same (Jacobi) view order, bit for bit.  Checked on the first
    # round of the schedule (its passes run at the coarsest level: cheap) over ALL views: rank 0 holds every depth map after
    # the exchange and compares them with its own single-context run.
    job_check = None
    if world > 1 and not args.no_job_check:
        nchk = 1 + args.geom_iters
        step(Timing(), passes=nchk)
        if rank == 0:
            got = [ctx.view_download(v)[0] for v in range(V)]
            solo = Context(local_rank)
            upload(solo)
            s1 = default_schedule()
            s1.rounds, s1.geom_iterations, s1.seed, s1.jacobi = args.rounds, args.geom_iters, 1, 1
            for p in range(nchk):
                solo.run_schedule_pass(s1, p, Timing())
            same = [bool(np.array_equal(got[v], solo.view_download(v)[0])) for v in range(V)]
            solo.close()
            job_check = {"passes": nchk, "views": V, "depth_maps_identical_to_single_gpu_jacobi": int(sum(same)), "ok": all(same)}
            if not all(same):
                sys.stderr.write("bench: the %d-GPU job differs from the single-GPU Jacobi run on views %s\n" % (world, [v for v in range(V) if not same[v]]))
        upload()
        barrier()

    for _ in range(args.warmup):
        step(Timing())
    # ---- timed region 1: device-resident inputs
    ctx.counters(reset=True)
    ctx.stage_stats(reset=True)
    ctx.set_profiling(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    tm = Timing()
    for _ in range(args.steps):
        step(tm)
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    ctx.set_profiling(False)
    st_ms, st_launch, st_evals = ctx.stage_stats()
    cnt = ctx.counters()
    anchor_patches = ctx.anchor_evals()
    # device time of the steps: CUDA events on the launching stream around every pass, including the wait for the exchange
    dev_s = tm.device_ms * 1e-3
    # ---- timed region 2: end to end from host buffers
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        upload()
        step(Timing())
        for v in range(first, first + count):
            d2h += sum(a.nbytes for a in ctx.view_download(v))
    barrier()
    e2e_s = time.perf_counter() - t0

    step_s, wall_s, e2e_s, exch_ms = reduce_max([dev_s, wall, e2e_s, tm.exchange_ms])
    evals_old, evals_new, launches_total, d2h_total, exch_bytes, cross_total = reduce_sum(
        [float(cnt[0]), float(cnt[1]), float(cnt[3]), float(d2h), float(tm.exchange_bytes), float(cross)])
    evals_total = evals_old + evals_new
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    value = V * args.steps / step_s
    fp32_peak, tex_peak = ctx.microbench()
    k = int(np.argmax(st_ms))
    k_evals = float(st_evals[k, 0] + st_evals[k, 1])
    k_ms = float(st_ms[k]) / max(1, int(st_launch[k]))
    k_ev_per_launch = k_evals / max(1, int(st_launch[k]))
    fp32_ach = k_ev_per_launch * ALG_FLOP_PER_EVAL / (k_ms * 1e-3) / 1e12
    tex_ach = k_ev_per_launch * SAMPLES_PER_EVAL / (k_ms * 1e-3) / 1e9
    frac_fp32, frac_tex = fp32_ach / fp32_peak, tex_ach / tex_peak
    headline_shape = (args.width, args.height, args.src) == (1920, 1080, 10)
    roofline = {
        "kernel": NCU_TRAFFIC.get(STAGE_NAMES[k], (None, "k_" + STAGE_NAMES[k], ""))[1], "stage": STAGE_NAMES[k],
        "share_of_step": float(st_ms[k] / st_ms.sum()),
        "bound": "l1tex" if frac_tex >= frac_fp32 else "fp32",
        "achieved": tex_ach if frac_tex >= frac_fp32 else fp32_ach,
        "peak": tex_peak if frac_tex >= frac_fp32 else fp32_peak,
        "unit": "Gsample/s" if frac_tex >= frac_fp32 else "TFLOP/s",
        "frac": max(frac_tex, frac_fp32),
        "traffic": NCU_TRAFFIC[STAGE_NAMES[k]][0] if headline_shape and STAGE_NAMES[k] in NCU_TRAFFIC else None,
        "traffic_unit": "DRAM bytes per full-resolution launch (ncu --set full, %s)" % NCU_TRAFFIC.get(STAGE_NAMES[k], (0, "", "n/a"))[2],
        "fp32": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak, "frac": frac_fp32, "flop_per_eval": ALG_FLOP_PER_EVAL},
        "l1tex": {"achieved_gsamples": tex_ach, "peak_gsamples": tex_peak, "frac": frac_tex, "samples_per_eval": SAMPLES_PER_EVAL,
                  "achieved_GBps": tex_ach * ALG_L1_BYTES_PER_EVAL / SAMPLES_PER_EVAL},
        "evals_per_launch": k_ev_per_launch, "ms_per_launch": k_ms,
        "peak_source": "measured in-process (apde_microbench): FMA chain / 6x6 bilinear gather; MEASURED_PEAKS.json holds no "
                       "texture-filter or FP32 peak (HBM / BF16 only)",
        "stage_ms": {STAGE_NAMES[i]: float(st_ms[i]) for i in range(11) if st_launch[i]},
        # per heavy stage: fraction of the measured texture-gather peak, 36 samples per evaluation (deformable evaluations
        # gather 36 + 9 per valid anchor; they are counted at 36 here, prop_weak_anchor_frac below adds the anchors' worst case)
        "stage_tex_frac": {STAGE_NAMES[i]: float((st_evals[i, 0] + st_evals[i, 1]) * SAMPLES_PER_EVAL / (st_ms[i] * 1e-3) / 1e9 / tex_peak)
                           for i in range(11) if st_launch[i] and st_ms[i] > 0 and (st_evals[i, 0] + st_evals[i, 1]) > 0},
        "stage_gevals": {STAGE_NAMES[i]: float((st_evals[i, 0] + st_evals[i, 1]) / 1e9) for i in range(11) if st_launch[i]},
    }
    i_pw = STAGE_NAMES.index("prop_weak")
    if st_launch[i_pw] and st_ms[i_pw] > 0:
        # the deformable cost's real sample count on this rank: 36 per centre patch + 9 per anchor patch actually sampled (counted
        # on the device by the anchor-sorted pipeline); the weak stage is the only one with deformable evaluations after init
        roofline["prop_weak_tex_frac_counted_anchors"] = float(
            (st_evals[i_pw, 0] * SAMPLES_PER_EVAL + st_evals[i_pw, 1] * SAMPLES_PER_EVAL + anchor_patches * SAMPLES_PER_ANCHOR) / (st_ms[i_pw] * 1e-3) / 1e9 / tex_peak)
        roofline["anchor_patches_per_deformable_eval"] = float(anchor_patches / max(1.0, float(st_evals[i_pw, 1])))
        roofline["prop_weak_tex_frac_with_8_anchors"] = float(
            (st_evals[i_pw, 0] * SAMPLES_PER_EVAL + st_evals[i_pw, 1] * (SAMPLES_PER_EVAL + 8 * SAMPLES_PER_ANCHOR)) / (st_ms[i_pw] * 1e-3) / 1e9 / tex_peak)
    per_arc, arcs = scene_layout(args, world)
    out = {
        "metric": "ref-views/s at %dx%d, %d src views" % (args.width, args.height, args.src),
        "value": value, "unit": "ref-views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: synthetic office scene, %d views (%d arcs x %d) %dx%d, %d src views, reference schedule: %d passes/view "
                               "(rounds x (1 photometric + %d geometric))" % (args.config, V, arcs, per_arc, args.width, args.height, args.src, npass, args.geom_iters),
                   "views_total": V, "passes_per_view": npass, "l2": "inputs larger than L2 (%.0f MB of images + maps per step)"
                   % (V * args.width * args.height * 26 / 1e6),
                   "weak_texture_share": args.weak,
                   "ordering": "reference (Gauss-Seidel)" if world == 1 else
                               "one scene over %d GPUs: views dealt out in blocks, Jacobi order, depth maps broadcast over NCCL view by view" % world},
        "cost_evals_per_s": evals_total / step_s, "cost_evals_per_step": evals_total / args.steps,
        "patchmatch_ms_per_view_pass": tm.patchmatch_ms / (args.steps * npass * max(count, 1)),
        "wall_ms_per_step": wall_s / args.steps * 1e3,
        "roofline": roofline,
        "e2e": {"value": V * args.steps / e2e_s, "unit": "ref-views/s",
                "h2d_bytes_per_step": int(world * (V * args.width * args.height + V * 120)), "d2h_bytes_per_step": int(d2h_total // args.steps)},
        "gpu_launches": int(launches_total),
        "clocks": clk,
    }
    if world > 1:
        out["exchange"] = {"exposed_ms_per_pass": exch_ms / (args.steps * npass), "bytes_per_pass_per_rank": exch_bytes / world / (args.steps * npass),
                           "source_reads_crossing_ranks": int(cross_total), "source_reads_total": int(V * args.src),
                           "collective": "ncclBroadcast per finished view (grouped over the ranks) on a second stream, overlapped with the next view",
                           "job_check": job_check}
    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline(args, scene, evals_total / args.steps / V)
    if world == 1 and not args.no_fusion:
        out["fusion"] = fusion_report(args, ctx, scene, cpu=not args.no_cpu_baseline)
    print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- fusion (outside the metric)
def fusion_report(args, ctx, scene, cpu=True):
    """WeakVisFilter + RunFusion of the maps the timed steps left on the device: GPU wall time for the whole scene, and the
    reference's host-side fusion (APD.cpp:962-1227, restated in the oracle, one thread as in the reference's main loop) on a
    bounded sample -- the first 4 views with their neighbours among those -- timed beside the GPU on the same sample."""
    from apde_mvs_b200.binding import Context
    from apde_mvs_b200.scene import Scene
    V = len(scene.images)
    ctx.fuse(True)  # warm-up: the first call allocates the skip / claim / candidate buffers
    t0 = time.perf_counter()
    xyz, _ = ctx.fuse(True)
    rep = {"gpu_ms_scene": (time.perf_counter() - t0) * 1e3, "points_scene": int(len(xyz)), "views_scene": V}
    nv = min(4, V)
    if nv < 3:
        return rep
    sub = Scene(scene.width, scene.height)
    sub.K = scene.K
    sub.images, sub.cameras = scene.images[:nv], scene.cameras[:nv]
    sub.pairs = [[q for q in range(nv) if q != v] for v in range(nv)]
    maps = [ctx.view_download(v) for v in range(nv)]
    c2 = Context(ctx.device)
    c2.load_scene(sub)
    for v in range(nv):
        c2.view_upload(v, *maps[v])
    c2.fuse(True)  # warm-up (allocations)
    t0 = time.perf_counter()
    gx, _ = c2.fuse(True)
    rep.update({"sample": "views 0..%d, %d neighbours each, %dx%d" % (nv - 1, nv - 1, scene.width, scene.height),
                "gpu_ms_sample": (time.perf_counter() - t0) * 1e3, "points_sample": int(len(gx))})
    c2.close()
    if cpu:
        from oracle import binding as orc
        depths, normals, weaks, confs = (np.stack([m[i] for m in maps]) for i in range(4))
        t0 = time.perf_counter()
        ox, _, _ = orc.fusion(sub.cameras, depths, normals, weaks, confs, sub.pairs, None, num_threads=os.cpu_count() or 1)
        rep.update({"cpu_ms_sample": (time.perf_counter() - t0) * 1e3, "cpu_kind": "port", "cpu_threads_weak_vis_filter": os.cpu_count() or 1,
                    "cpu_threads_fusion": 1, "identical_to_cpu": bool(len(ox) == len(gx) and np.allclose(ox, gx, rtol=1e-5, atol=1e-5))})
    return rep


# ----------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(args, scene, evals_per_view):
    """the CPU oracle (a port: the reference has no CPU PatchMatch) on a bounded sample: one photometric pass of one view
    at the coarsest pyramid level, all host threads.  Projected to ref-views/s through the measured evals per view."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import ref_params
    from oracle import binding as orc
    cores = os.cpu_count() or 1
    w, h = max(64, scene.width // 4), max(48, scene.height // 4)  # ~10 s of CPU work at 1920x1080 x 10 views on 16 cores
    ids = [0] + list(scene.pairs[0])
    imgs, cams = [], []
    for i in ids:
        imgs.append(cv2.resize(scene.images[i].astype(np.float32), (w, h), interpolation=cv2.INTER_LINEAR))
        cam = type(scene.cameras[i])()
        C.memmove(C.byref(cam), C.byref(scene.cameras[i]), C.sizeof(cam))
        sx, sy = w / float(scene.width), h / float(scene.height)
        cam.K[0] *= sx; cam.K[2] *= sx; cam.K[4] *= sy; cam.K[5] *= sy
        cam.width, cam.height = w, h
        cams.append(cam)
    p = ref_params()
    p.depth_min, p.depth_max = cams[0].depth_min * 0.6, cams[0].depth_max * 1.2
    pb = orc.Problem(imgs, cams, p, seed=1, stream=0, tex_mode=1, num_threads=cores)
    t0 = time.perf_counter()
    pb.stage("run_pass")
    dt = time.perf_counter() - t0
    ev = float(sum(pb.counters()[:2]))
    return {"value": (ev / dt) / evals_per_view, "unit": "ref-views/s (projected: oracle cost evals/s / evals per view)",
            "cores": cores, "kind": "port", "evals_per_s": ev / dt,
            "sample": "one photometric pass, view 0, %dx%d, %d src views, %.1f s" % (w, h, len(ids) - 1, dt)}


# ----------------------------------------------------------------------------------------------- reference arm
def reference_main_run(args, scene, tmp):
    """ONE run of the reference's own, unmodified main() (APD.cu + APD.cpp + main.cpp in oracle/_ref/libapd_ref_full.so) over a
    dense folder holding the scene: its GenerateSampleList, InuputInitialization (pyramid, map hand-over), ProcessProblem,
    schedule and .bin output, with --memory_cache true --no_fuse true.  Returns views/s end to end (wall clock of main()) and
    views/s of its own 'RunPatchMatch time' hook (main.cpp:157-161), or None when the library is not there."""
    from oracle import ref_main_runner as runner
    if not runner.available():
        return None
    scene.write_dense_folder(tmp)
    V = len(scene.images)
    for v in range(V):  # PGM content under the .png name the reference looks for (its stub cv::imread decodes by magic)
        with open(os.path.join(tmp, "images", "%08d.png" % v), "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (scene.width, scene.height))
            f.write(np.ascontiguousarray(scene.images[v]).tobytes())
    t0 = time.perf_counter()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_main_runner.py"), "4242", "--dense_folder", tmp, "--use_sa", "false",
                          "--memory_cache", "true", "--no_fuse", "true", "--dataset", "General"], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if out.returncode != 0:
        return {"error": (out.stdout[-300:] + out.stderr[-300:]).replace("\n", " ")}
    pm = sum(int(ln.split(":")[1].split("ms")[0]) for ln in out.stdout.splitlines() if ln.startswith("RunPatchMatch time:"))
    cost = [int(ln.split(":")[1].split("ms")[0]) for ln in out.stdout.splitlines() if ln.startswith("Cost time:")]
    return {"e2e_views_per_s": V / wall, "wall_s": wall, "patchmatch_views_per_s": V / (pm * 1e-3) if pm else None,
            "main_cost_time_ms": cost[-1] if cost else None, "views": V,
            "what": "one run of the reference's unmodified main() on a dense folder (--memory_cache true --no_fuse true), process start to exit"}


def run_reference(args, rank, world):
    """The reference's own code on the same scene and schedule.  value: the reference's APD.cu (oracle/_ref/libapd_ref.so,
    compiled unmodified for sm_100 with the reference flags) through the schedule of main.cpp:303-367, K steps, each step a
    bounded sample of the workload (--ref-views reference views through all passes; every view's first photometric pass runs
    because the later passes read those depth maps), timed by the reference's own 'RunPatchMatch time' hook.  e2e: one run of
    the reference's unmodified main() (libapd_ref_full.so) over the whole scene."""
    if rank != 0:
        return
    from oracle import ref_binding as ref
    from oracle.ref_schedule import compute_round_num, run_reference_schedule, sample_views
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so not built (needs /root/reference at build time)"}))
        return
    scene = build_scene(args, 1)
    V = len(scene.images)
    nviews = len(sample_views(V, args.ref_views))
    W, H = scene.width, scene.height
    rounds = args.rounds if args.rounds > 0 else compute_round_num(W, H)

    def one_step():
        return run_reference_schedule(scene, rounds, args.geom_iters, nviews)[1]

    for _ in range(args.warmup):
        one_step()
    pm = 0.0
    for _ in range(args.steps):
        pm += one_step()
    value = nviews * args.steps / (pm * 1e-3)
    import tempfile
    with tempfile.TemporaryDirectory(prefix="apde_ref_main_") as tmp:
        main_run = reference_main_run(args, scene, tmp)
    e2e_value = main_run["e2e_views_per_s"] if main_run and "e2e_views_per_s" in main_run else None
    print(json.dumps({
        "impl": "reference", "metric": "ref-views/s at %dx%d, %d src views" % (W, H, args.src), "value": value,
        "unit": "ref-views/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": pm / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: same scene and schedule as the default arm; the reference's APD.cu rebuilt for sm_100 (reference flags) "
                               "run on the GPU; %d of %d reference views timed per step" % (args.config, nviews, V),
                   "views_total": V, "views_timed": sample_views(V, args.ref_views), "passes_per_view": rounds * (1 + args.geom_iters)},
        "cpu_baseline": {"value": value, "unit": "ref-views/s", "kind": "reference", "cores": 1,
                         "sample": "%d of %d reference views through all %d passes per step (plus the first photometric pass of every view); "
                                   "sum of the reference's own 'RunPatchMatch time' (main.cpp:157-161)" % (nviews, V, rounds * (1 + args.geom_iters))},
        "e2e": {"value": e2e_value if e2e_value is not None else value, "unit": "ref-views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "source": "reference main()" if e2e_value is not None else "unavailable: falls back to value"},
        "reference_main": main_run,
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched by torchrun: re-launch one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
