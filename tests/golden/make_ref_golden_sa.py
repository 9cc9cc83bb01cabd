"""Generates tests/golden/ref_costs_sa.npz ON THE GPU BOX: the inputs of ref_costs.npz (same images, cameras, tuples, planes,
weak map, anchors) evaluated by the reference's own ComputeBilateralNCCOld / ComputeBilateralNCCNew WITH a segment-label map
(sa_mask_cuda, APD.cpp:641-649, 764-766) -- NCC-Old branch B (APD.cu:664-719) and the label tests of NCC-New.

    gpurun -- 'python tests/golden/make_ref_golden_sa.py'   ->  gpurun_out/ref_costs_sa.npz  (copied to tests/golden/)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import ref_params  # noqa: E402
from oracle import binding as orc  # noqa: E402
from oracle import ref_binding as ref  # noqa: E402
from apde_mvs_b200.scene import make_label_map as make_labels  # noqa: E402


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_costs.npz"))
    imgs = list(z["images"])
    cams = []
    for row in z["cams"]:
        cam = orc.OCamera()
        C.memmove(C.byref(cam), row.tobytes(), 120)
        cams.append(cam)
    h, w = imgs[0].shape
    p = ref_params()
    p.depth_min, p.depth_max = float(z["iparams"][0]), float(z["iparams"][1])
    lab = make_labels(w, h, 5, zero_share=0.2)
    out = {"labels": lab}
    ref.set_sa_mask(lab)
    out["old_costs"] = ref.eval_costs(imgs, cams, p, z["old_tuples"], z["old_planes"], 0)
    p.use_APD = 1
    out["new_costs"] = ref.eval_costs(imgs, cams, p, z["new_tuples"], z["new_planes"], 1, weak=z["new_weak"], selected_views=z["new_sel"],
                                      anchors=z["new_anchors"])
    ref.set_sa_mask(None)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_costs_sa.npz"), **out)
    print("labels: %d segments, %.2f unlabelled; old costs changed by the map %.3f, new costs changed %.3f" % (
        len(np.unique(lab)), (lab == 0).mean(), (np.abs(out["old_costs"] - z["old_costs"]) > 1e-3).mean(),
        (np.abs(out["new_costs"] - z["new_costs"]) > 1e-3).mean()))


if __name__ == "__main__":
    main()
