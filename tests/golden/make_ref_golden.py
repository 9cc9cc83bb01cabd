"""Generates tests/golden/ref_costs.npz ON THE GPU BOX from the reference's own device functions
(oracle/_ref/libapd_ref.so == /root/reference/APD.cu compiled unmodified, see oracle/ref_wrapper.cu).

    gpurun -- 'python tests/golden/make_ref_golden.py'   ->  gpurun_out/ref_costs.npz  (copied to tests/golden/)

The file pins the CPU oracle (tests/test_oracle_cpu.py): NCC-Old, NCC-New (deformable) and geometric-consistency
costs on seeded (pixel, view, plane) tuples of a small synthetic scene with weak-texture blobs, plus one whole
reference pass (RunPatchMatch, seed patched) for the statistical depth check.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import ref_params  # noqa: E402
from apde_mvs_b200.scene import make_office_scene  # noqa: E402
from oracle import binding as orc  # noqa: E402
from oracle import ref_binding as ref  # noqa: E402


def cam_bytes(cams):
    return np.frombuffer(b"".join(bytes(C.string_at(C.byref(c), 120)) for c in cams), np.uint8).reshape(len(cams), 120).copy()


def main():
    scene = make_office_scene(128, 96, num_views=5, num_src=4, seed=9, weak=0.3)
    ref_view = 2
    ids = [ref_view] + list(scene.pairs[ref_view])
    imgs = [scene.images[i].astype(np.float32) for i in ids]
    cams = [scene.cameras[i] for i in ids]
    h, w = imgs[0].shape
    rng = np.random.default_rng(2024)
    p = ref_params()
    p.depth_min, p.depth_max = cams[0].depth_min * 0.6, cams[0].depth_max * 1.2
    out = {"images": np.stack(imgs), "cams": cam_bytes(cams)}

    # ---- one whole reference pass (photometric, FIRST_INIT) -> planes / depth for the statistical check and as state
    planes, weak, conf, ms = ref.run_pass(imgs, cams, p, seed=4242)
    out["pass_planes"], out["pass_weak"] = planes, weak
    depth = planes[..., 3]
    gt = scene.gt_depth[ref_view]
    out["gt_depth"] = gt

    # ---- NCC-Old: random + near-truth planes (camera-frame normal, plane distance)
    N = 6000
    xs, ys, vs = rng.integers(0, w, N), rng.integers(0, h, N), rng.integers(1, len(ids), N)
    fx, fy, cx, cy = cams[0].K[0], cams[0].K[4], cams[0].K[2], cams[0].K[5]
    R = np.array(cams[0].R, np.float64).reshape(3, 3)
    nw = planes[ys, xs, :3].astype(np.float64)
    nc = (R @ nw.T).T
    d = np.where(gt[ys, xs] > 0, gt[ys, xs], depth[ys, xs]) * rng.uniform(0.97, 1.03, N)
    X = np.stack([d * (xs - cx) / fx, d * (ys - cy) / fy, d], 1)
    pl = np.concatenate([nc, -(np.sum(nc * X, 1))[:, None]], 1).astype(np.float32)
    t = np.stack([xs, ys, vs], 1).astype(np.int32)
    out["old_tuples"], out["old_planes"] = t, pl
    out["old_costs"] = ref.eval_costs(imgs, cams, p, t, pl, 0)

    # ---- geometric cost: source depth maps = the ground truth (0 where nothing is hit)
    depths = [scene.gt_depth[i].astype(np.float32) for i in ids]
    out["depths"] = np.stack(depths)
    out["geom_costs"] = ref.eval_costs(imgs, cams, p, t, pl, 2, depths=depths)

    # ---- NCC-New: weak map from the reference pass, anchors from the oracle's GenAnchors (inputs, not under test here)
    po = orc.OParams()
    C.memmove(C.byref(po), C.byref(p), C.sizeof(po))
    po.use_APD, po.rotate_time, po.ransac_threshold, po.state = 1, 2, 0.00875, 1
    pb = orc.Problem(imgs, cams, po, depths=depths, seed=7, stream=ref_view, tex_mode=1)
    pb.planes[...] = planes
    pb.weak_info[...] = weak
    pb.confidence[...] = 1
    pb.stage("nearest_strong")
    pb.stage("gen_anchors")
    pb.stage("neighbour_update")
    wy, wx = np.nonzero(pb.weak_info == 0)
    print("weak pixels with anchors:", len(wx))
    pick = rng.choice(len(wx), min(4000, len(wx)), replace=False)
    t2 = np.stack([wx[pick], wy[pick], rng.integers(1, len(ids), len(pick))], 1).astype(np.int32)
    nw = planes[wy[pick], wx[pick], :3].astype(np.float64)
    nc = (R @ nw.T).T
    d = planes[wy[pick], wx[pick], 3].astype(np.float64) * rng.uniform(0.98, 1.02, len(pick))
    X = np.stack([d * (t2[:, 0] - cx) / fx, d * (t2[:, 1] - cy) / fy, d], 1)
    pl2 = np.concatenate([nc, -(np.sum(nc * X, 1))[:, None]], 1).astype(np.float32)
    sel = rng.integers(0, 16, (h, w)).astype(np.uint32)
    out["new_tuples"], out["new_planes"], out["new_weak"], out["new_anchors"], out["new_sel"] = t2, pl2, pb.weak_info.copy(), pb.anchors.copy(), sel
    p.use_APD = 1
    out["new_costs"] = ref.eval_costs(imgs, cams, p, t2, pl2, 1, weak=pb.weak_info, selected_views=sel, anchors=pb.anchors)
    out["iparams"] = np.array([p.depth_min, p.depth_max], np.float32)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_costs.npz"), **out)
    print("reference pass: %.1f ms; cost hist old %s new %s geom %s" % (
        ms, np.histogram(out["old_costs"], bins=[0, .2, 1, 1.999, 2])[0], np.histogram(out["new_costs"], bins=[0, .2, 1, 1.999, 2])[0],
        np.histogram(out["geom_costs"], bins=[0, .5, 2.999, 3])[0]))


if __name__ == "__main__":
    main()
