"""The oracle's HOST-side restatement pinned against the reference's own host code: /root/reference/APD.cpp compiled
unmodified into oracle/_ref/libapd_ref_host.so (make -C oracle ref_host) and run here on the CPU.  Covers ReadCamera,
the .bin map format, WeakVisFilter + RunFusion and the two Tanks-and-Temples fusion variants, PLY export."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import ROOT
from oracle import binding as orc
from oracle import ref_host_binding as refh

pytestmark = pytest.mark.skipif(not refh.available(), reason="oracle/_ref/libapd_ref_host.so not built (needs /root/reference at build time)")


def _write_bin(path, a):
    """WriteBinMat layout (APD.cpp:58-83): int32 version = 1, rows, cols, cv type, then the rows"""
    a = np.ascontiguousarray(a)
    typ = {(np.dtype(np.uint8), 1): 0, (np.dtype(np.float32), 1): 5, (np.dtype(np.float32), 3): 21}[(a.dtype, 1 if a.ndim == 2 else a.shape[2])]
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 1, a.shape[0], a.shape[1], typ))
        f.write(a.tobytes())


def _dense_folder(tmp_path, W=112, H=80, V=5, N=4, seed=11):
    from apde_mvs_b200.scene import make_office_scene
    scene = make_office_scene(W, H, num_views=V, num_src=N, seed=7, arc_deg=20.0, with_color=True)
    d = tmp_path / "scan"
    scene.write_dense_folder(str(d))
    rng = np.random.default_rng(seed)
    depths = np.stack(scene.gt_depth).astype(np.float32)
    depths *= (1 + rng.normal(0, 3e-4, depths.shape)).astype(np.float32)  # noise around the fusion thresholds
    out = rng.random(depths.shape) < 0.10
    depths[out] *= rng.uniform(0.85, 0.97, out.sum()).astype(np.float32)  # floaters in front of the surfaces: WeakVisFilter's prey
    depths[rng.random(depths.shape) < 0.12] = 0
    depths[:, 30:42, 20:70] = 0
    normals = rng.normal(0, 1, (V, H, W, 3)).astype(np.float32) * 0.05 + np.array([0, 0, -1], np.float32)
    normals /= np.linalg.norm(normals, axis=-1, keepdims=True)
    weaks = (rng.random((V, H, W)) < 0.7).astype(np.uint8)
    weaks[rng.random((V, H, W)) < 0.05] = 2
    confs = rng.integers(0, 255, (V, H, W)).astype(np.uint8)
    for v in range(V):
        r = d / "APD" / ("%08d" % v)
        os.makedirs(r)
        _write_bin(r / "depths.bin", depths[v]); _write_bin(r / "normals.bin", normals[v])
        _write_bin(r / "weak.bin", weaks[v]); _write_bin(r / "confidence.bin", confs[v])
        rgb = scene.colors[v][..., ::-1]  # scene colours are BGR; PPM stores RGB
        with open(d / "images" / ("%08d.ppm" % v), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (W, H))
            f.write(np.ascontiguousarray(rgb).tobytes())
    return d, scene, depths, normals, weaks, confs


def test_read_camera_and_bin_format_match_the_reference(tmp_path):
    d, scene, depths, normals, weaks, confs = _dense_folder(tmp_path)
    # ReadCamera: the reference's parse of the files == the cameras the product / oracle use (c = -R^T t in double, APD.cpp:114-119)
    from apde_mvs_b200 import build as b
    b.build_host()
    tool = os.path.join(ROOT, "apde_mvs_b200", "_build", "test_io")
    for v in range(len(scene.images)):
        cam = refh.read_camera(d / "cams" / ("%08d_cam.txt" % v))
        mine = np.array(subprocess.check_output([tool, "cam", str(d / "cams" / ("%08d_cam.txt" % v))]).split(), np.float64)
        ref = np.concatenate([cam["K"], cam["R"], cam["t"], cam["c"], [cam["depth_min"], cam["depth_max"], cam["interval"], cam["depth_num"]]])
        assert np.allclose(mine, ref, rtol=1e-6, atol=1e-6), (v, mine, ref)
        sc = scene.cameras[v]
        assert np.allclose(np.array(sc.K[:]), cam["K"], rtol=1e-6) and np.allclose(np.array(sc.c[:]), cam["c"], rtol=1e-5, atol=1e-5)
    # the .bin format: files written here are read and re-written by the reference byte for byte, and read by the product
    for name, typ in (("depths.bin", 5), ("normals.bin", 21), ("weak.bin", 0), ("confidence.bin", 0)):
        src = d / "APD" / "00000001" / name
        assert refh.copy_bin_mat(src, tmp_path / "copy.bin") == typ
        assert open(src, "rb").read() == open(tmp_path / "copy.bin", "rb").read()
        out = subprocess.check_output([tool, "bin", str(tmp_path / "copy.bin"), str(tmp_path / "copy2.bin")]).split()
        assert int(out[2]) == typ and open(tmp_path / "copy2.bin", "rb").read() == open(src, "rb").read()


def test_sample_list_and_round_rule_match_the_reference(tmp_path):
    """pair.txt parsing (score <= 0 dropped, extension probing) and ComputeRoundNum: the product's C++ host side and the
    schedule's round rule against the reference's own GenerateSampleList / ComputeRoundNum (main.cpp:44-146)"""
    from apde_mvs_b200 import build as b
    b.build_host()
    tool = os.path.join(ROOT, "apde_mvs_b200", "_build", "test_io")
    from oracle.ref_schedule import compute_round_num
    for (w, h), want in (((640, 480), 1), ((801, 600), 2), ((1600, 1200), 2), ((1920, 1056), 3), ((400, 3300), 4)):
        d = tmp_path / ("s_%dx%d" % (w, h))
        os.makedirs(d / "images"); os.makedirs(d / "cams"); os.makedirs(d / "APD")  # main() creates APD/ before the list (main.cpp:262)
        for v in range(3):
            with open(d / "images" / ("%08d.png" % v), "wb") as f:   # PGM content under the name the reference probes for
                f.write(b"P5\n%d %d\n255\n" % (w, h)); f.write(bytes(w * h))
        with open(d / "pair.txt", "w") as f:
            f.write("3\n0\n3 1 10.5 2 0.0 1 -1\n1\n2 0 1 2 3.5\n2\n0\n")
        ids, src, ext = refh.generate_sample_list(d)
        assert ids == [0, 1, 2] and src == [[1], [0, 2], []] and ext == ".png"
        assert refh.compute_round_num(d) == want == compute_round_num(w, h)
        out = subprocess.check_output([tool, "pairs", str(d)], text=True).strip().splitlines()
        mine = [(int(l.split()[0]), [int(x) for x in l.split(":")[1].split()]) for l in out]
        assert [m[0] for m in mine] == ids and [m[1] for m in mine] == src and out[0].split()[1].rstrip(":") == ext


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_fusion_oracle_matches_the_reference_code(tmp_path, variant):
    """WeakVisFilter + RunFusion / RunFusion_TAT_I / RunFusion_TAT_A: the oracle reproduces the reference's point cloud
    point for point, in order, with its colours, with and without the weak filter"""
    d, scene, depths, normals, weaks, confs = _dense_folder(tmp_path)
    V = len(scene.images)
    cams = []
    for v in range(V):  # the cameras exactly as the reference parses them
        c = refh.read_camera(d / "cams" / ("%08d_cam.txt" % v))
        oc = orc.OCamera()
        for i in range(9):
            oc.K[i], oc.R[i] = float(c["K"][i]), float(c["R"][i])
        for i in range(3):
            oc.t[i], oc.c[i] = float(c["t"][i]), float(c["c"][i])
        oc.height, oc.width = c["height"], c["width"]
        oc.depth_min, oc.depth_max, oc.interval, oc.depth_num = c["depth_min"], c["depth_max"], c["interval"], c["depth_num"]
        cams.append(oc)
    for wf in (False, True):
        xyz_r, bgr_r = refh.run_fusion(d, scene.pairs, variant=variant, weak_filter=wf)
        xyz_o, bgr_o, _ = orc.fusion(cams, depths, normals, weaks, confs, scene.pairs, np.stack(scene.colors), weak_filter=wf, variant=variant)
        print("variant %d weak_filter %d: reference %d points, oracle %d" % (variant, wf, len(xyz_r), len(xyz_o)))
        assert len(xyz_r) > 300
        assert len(xyz_o) == len(xyz_r)
        assert np.array_equal(xyz_o, xyz_r)
        assert np.array_equal(bgr_o.astype(np.uint8), bgr_r)  # ExportPointCloud stores (uchar) colour, APD.cpp:340-346
    # WeakVisFilter itself: the reference's skip.png (APD.cpp:1026-1036, stored as PGM by the stub imwrite) == the oracle's skip maps
    _, _, skip = orc.fusion(cams, depths, normals, weaks, confs, scene.pairs, np.stack(scene.colors), weak_filter=True, variant=variant)
    assert skip.sum() > 50
    for v in range(V):
        raw = open(d / "APD" / ("%08d" % v) / "skip.png", "rb").read().split(b"\n", 3)
        assert raw[0] == b"P5"
        ref_skip = np.frombuffer(raw[3], np.uint8).reshape(skip[v].shape)
        assert np.array_equal(ref_skip == 255, skip[v] == 1), "skip map of view %d differs from the reference's" % v
