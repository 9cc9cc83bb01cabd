#!/usr/bin/env python
"""bench.py -- ref-views/s and cost evals/s of the APDe-MVS PatchMatch hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo (libapde.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own APD.cu (oracle/_ref) on the same scene

One "step" = the whole multi-scale schedule of main.cpp:303-367 (R rounds x (1 photometric + 3 geometric) passes) over
every reference view of a synthetic 1920x1080 scene with 10 source views per reference view.  `value` = views completed
per second of device time (inputs resident in HBM); `e2e` = the same through the C ABI from pinned HOST buffers
(scene upload + schedule + depth-map download inside the timed region).  N > 1: views are sharded over the ranks (weak
scaling: 11 views per GPU), depth maps are all-gathered over NCCL between passes (Jacobi ordering).
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
# dram__bytes_read.sum + dram__bytes_write.sum of one full-resolution k_prop_strong launch at the default workload
# (1920x1080, 10 source views), from the `ncu --set full` capture summarised in
# profiles/r01_ncu_full_prop_strong_compacted.md: 193-209 MB read + 55-61 MB written.  The algorithmic minimum of that
# launch is ~115 MB (11 u8 images, plane / cost / mask reads of the frame, writes of the half frame): the kernel is bound by
# the texture data pipe (88-89 % of its wavefront peak), not by HBM (0.7 % of DRAM throughput).
NCU_DRAM_BYTES_PROP_STRONG_1080P = 2.6e8
# k_sweep_columns (DepthToWeak), same workload, profiles/r01_ncu_1080p_raw_summary.txt: 0.56 GB read + 2.99 GB written per
# full-resolution launch -- the [62][columns] cost arrays (algorithmic: columns x 62 x 4 B x 2 with the geometric term).
NCU_DRAM_BYTES_SWEEP_COLUMNS_1080P = 3.55e9
NCU_TRAFFIC = {"prop_strong": (NCU_DRAM_BYTES_PROP_STRONG_1080P, "k_prop_strong", "profiles/r01_ncu_full_prop_strong_compacted.md"),
               "depth_to_weak": (NCU_DRAM_BYTES_SWEEP_COLUMNS_1080P, "k_sweep_columns", "profiles/r01_ncu_1080p_raw_summary.txt")}
STAGE_NAMES = ["nearest_strong", "gen_anchors", "init", "prop_strong", "ransac_fit", "prop_weak", "depth_normal", "median",
               "depth_to_weak", "confidence", "local_refine"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--views-per-gpu", type=int, default=11)
    ap.add_argument("--src", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=0, help="0 = reference rule (ComputeRoundNum)")
    ap.add_argument("--geom-iters", type=int, default=3)
    ap.add_argument("--weak", type=float, default=0.0, help="share of every surface covered by weak-texture blobs (C3-like workloads; 0 = headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fusion", action="store_true", help="skip the fusion report (outside the timed metric)")
    ap.add_argument("--ref-views", type=int, default=0, help="reference arm: reference views processed per step (0 = all)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- scene
def build_scene(args, world):
    """base group of views_per_gpu views, replicated `world` times as independent groups (identical work per GPU)"""
    from apde_mvs_b200.scene import Scene, make_office_scene
    cache = "/tmp/apde_bench_scene_%dx%d_%d_%d_%g.npz" % (args.width, args.height, args.views_per_gpu, args.src, args.weak)
    pre = None
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            pre = list(zip(z["images"], z["gt"]))
        except Exception:
            pre = None
    base = make_office_scene(args.width, args.height, args.views_per_gpu, args.src, seed=2, weak=args.weak, prerendered=pre)
    if pre is None:
        try:
            np.savez(cache + ".tmp.npz", images=np.stack(base.images), gt=np.stack(base.gt_depth))
            os.replace(cache + ".tmp.npz", cache)
        except Exception:
            pass
    if world == 1:
        return base
    sc = Scene(args.width, args.height)
    sc.K = base.K
    nb = len(base.images)
    for g in range(world):
        sc.images += base.images
        sc.gt_depth += base.gt_depth
        sc.cameras += base.cameras
        sc.Rs += base.Rs
        sc.ts += base.ts
        sc.pairs += [[p + g * nb for p in pr] for pr in base.pairs]
    return sc


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
    """cudaHostRegister a numpy buffer so the C ABI's host->device copies come from pinned memory"""
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
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = build_scene(args, world)
    V = len(scene.images)
    vpg = args.views_per_gpu
    images = [pin(np.ascontiguousarray(im)) for im in scene.images]
    ctx = Context(local_rank)

    def upload():
        ctx.scene_begin(V, scene.width, scene.height)
        for v in range(V):
            ctx.scene_set_view(v, images[v], scene.cameras[v])
            ctx.scene_set_pairs(v, scene.pairs[v])
        ctx.scene_commit()

    sched = default_schedule()
    sched.rounds, sched.geom_iterations, sched.seed = args.rounds, args.geom_iters, 1
    if world > 1:
        sched.jacobi, sched.first_view, sched.num_views_local = 1, rank * vpg, vpg
    upload()
    npass = ctx.num_passes(sched)

    pool_tensor = {}

    def exchange():
        """all-gather the owned depth maps into every rank's replicated pool (in place, NCCL over NVLink)"""
        import torch
        ptr, nbytes, per_view = ctx.depth_pool()
        if ptr not in pool_tensor:
            class W:
                __cuda_array_interface__ = {"shape": (V, per_view // 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
            pool_tensor[ptr] = torch.as_tensor(W(), device="cuda:%d" % local_rank)
        t = pool_tensor[ptr]
        dist.all_gather_into_tensor(t, t[rank * vpg:(rank + 1) * vpg])
        torch.cuda.synchronize()

    def step(timing):
        for p in range(npass):
            ctx.run_schedule_pass(sched, p, timing)
            if world > 1:
                exchange()

    def barrier():
        if world > 1:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

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
    # device time of the steps: CUDA events on the launching stream (+ the exchange, which only wall clock sees)
    dev_s = tm.device_ms * 1e-3
    step_s = max(dev_s, 0.0) if world == 1 else wall
    # ---- timed region 2: end to end from host buffers
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        upload()
        step(Timing())
        for v in range(rank * vpg, (rank + 1) * vpg):
            d2h += ctx.view_download(v)[0].nbytes
    barrier()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        import torch
        t = torch.tensor([step_s, e2e_s], device="cuda:%d" % local_rank, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s, e2e_s = float(t[0]), float(t[1])
        ev = torch.tensor([float(cnt[0] + cnt[1]), float(cnt[3])], device="cuda:%d" % local_rank, dtype=torch.float64)
        dist.all_reduce(ev)
        evals_total, launches_total = float(ev[0]), float(ev[1])
    else:
        evals_total, launches_total = float(cnt[0] + cnt[1]), float(cnt[3])
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
    roofline = {
        "kernel": NCU_TRAFFIC.get(STAGE_NAMES[k], (None, "k_" + STAGE_NAMES[k], ""))[1], "stage": STAGE_NAMES[k],
        "share_of_step": float(st_ms[k] / st_ms.sum()),
        "bound": "l1tex" if frac_tex >= frac_fp32 else "fp32",
        "achieved": tex_ach if frac_tex >= frac_fp32 else fp32_ach,
        "peak": tex_peak if frac_tex >= frac_fp32 else fp32_peak,
        "unit": "Gsample/s" if frac_tex >= frac_fp32 else "TFLOP/s",
        "frac": max(frac_tex, frac_fp32),
        "traffic": NCU_TRAFFIC[STAGE_NAMES[k]][0] if (args.width, args.height, args.src) == (1920, 1080, 10) and STAGE_NAMES[k] in NCU_TRAFFIC else None,
        "traffic_unit": "DRAM bytes per full-resolution launch (ncu --set full, %s)" % NCU_TRAFFIC.get(STAGE_NAMES[k], (0, "", "n/a"))[2],
        "fp32": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak, "frac": frac_fp32, "flop_per_eval": ALG_FLOP_PER_EVAL},
        "l1tex": {"achieved_gsamples": tex_ach, "peak_gsamples": tex_peak, "frac": frac_tex, "samples_per_eval": SAMPLES_PER_EVAL,
                  "achieved_GBps": tex_ach * ALG_L1_BYTES_PER_EVAL / SAMPLES_PER_EVAL},
        "evals_per_launch": k_ev_per_launch, "ms_per_launch": k_ms,
        "peak_source": "measured in-process (apde_microbench): FMA chain / 6x6 bilinear gather",
        "stage_ms": {STAGE_NAMES[i]: float(st_ms[i]) for i in range(11) if st_launch[i]},
        # per heavy stage: fraction of the measured texture-gather peak (36 samples per NCC-Old evaluation; NCC-New
        # evaluations gather 36 + 9 per valid anchor, counted here at 36, so prop_weak is under-stated)
        "stage_tex_frac": {STAGE_NAMES[i]: float((st_evals[i, 0] + st_evals[i, 1]) * SAMPLES_PER_EVAL / (st_ms[i] * 1e-3) / 1e9 / tex_peak)
                           for i in range(11) if st_launch[i] and st_ms[i] > 0 and (st_evals[i, 0] + st_evals[i, 1]) > 0},
        "stage_gevals": {STAGE_NAMES[i]: float((st_evals[i, 0] + st_evals[i, 1]) / 1e9) for i in range(11) if st_launch[i]},
    }
    out = {
        "metric": "ref-views/s at %dx%d, %d src views" % (args.width, args.height, args.src),
        "value": value, "unit": "ref-views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic office scene, %d views/GPU %dx%d, %d src views, reference schedule: %d passes/view "
                               "(rounds x (1 photometric + %d geometric))" % (vpg, args.width, args.height, args.src, npass, args.geom_iters),
                   "views_total": V, "passes_per_view": npass, "l2": "inputs larger than L2 (%.0f MB of images + maps per step)"
                   % (V * args.width * args.height * 26 / 1e6),
                   "weak_texture_share": args.weak,
                   "ordering": "reference (Gauss-Seidel)" if world == 1 else "Jacobi + NCCL all-gather of depth maps per pass"},
        "cost_evals_per_s": evals_total / step_s, "cost_evals_per_step": evals_total / args.steps,
        "patchmatch_ms_per_view_pass": tm.patchmatch_ms / (args.steps * npass * vpg),
        "roofline": roofline,
        "e2e": {"value": V * args.steps / e2e_s, "unit": "ref-views/s",
                "h2d_bytes_per_step": int(V * args.width * args.height + V * 120), "d2h_bytes_per_step": int(d2h // args.steps)},
        "gpu_launches": int(launches_total),
        "clocks": clk,
    }
    if not args.no_cpu_baseline:
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
def run_reference(args, rank, world):
    """The reference's own APD.cu (oracle/_ref/libapd_ref.so, compiled unmodified for sm_100) driven through the same
    schedule as main.cpp:303-367; host-side stages of InuputInitialization / ProcessProblem are emulated with cv2/numpy."""
    if rank != 0:
        return
    from oracle import ref_binding as ref
    from oracle.ref_schedule import compute_round_num, run_reference_schedule
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so not built (needs /root/reference at build time)"}))
        return
    scene = build_scene(args, 1)
    V = len(scene.images)
    nviews = args.ref_views if args.ref_views > 0 else V
    W, H = scene.width, scene.height
    rounds = args.rounds if args.rounds > 0 else compute_round_num(W, H)

    def one_step():
        return run_reference_schedule(scene, rounds, args.geom_iters, nviews)[1]

    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    pm = 0.0
    for _ in range(args.steps):
        pm += one_step()
    wall = time.perf_counter() - t0
    value = nviews * args.steps / (pm * 1e-3)
    print(json.dumps({
        "impl": "reference", "metric": "ref-views/s at %dx%d, %d src views" % (W, H, args.src), "value": value,
        "unit": "ref-views/s", "n_gpus": 1, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": pm / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "same scene and schedule as the default arm; the reference's APD.cu rebuilt for sm_100 "
                               "(reference flags) and run on the GPU; %d of %d reference views timed per step" % (nviews, V),
                   "passes_per_view": rounds * (1 + args.geom_iters)},
        "cpu_baseline": {"value": value, "unit": "ref-views/s", "kind": "reference", "cores": 1,
                         "sample": "sum of the reference's own 'RunPatchMatch time' (main.cpp:157-161) over all passes; host "
                                   "stages (resize, map hand-over) emulated with cv2/numpy and excluded from value"},
        "e2e": {"value": nviews * args.steps / wall, "unit": "ref-views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
