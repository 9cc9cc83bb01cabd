"""GPU tests of the drop-in surface: the `apd` CLI (reference flag spellings, output files) and the C++ APD class."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _read_bin(p):
    with open(p, "rb") as f:
        version, rows, cols, typ = struct.unpack("<4i", f.read(16))
        assert version == 1
        dt, ch = {0: (np.uint8, 1), 5: (np.float32, 1), 21: (np.float32, 3)}[typ]
        a = np.frombuffer(f.read(), dt)
        return a.reshape(rows, cols, ch) if ch > 1 else a.reshape(rows, cols)


@pytest.fixture(scope="module")
def dense(tmp_path_factory):
    from apde_mvs_b200 import build as b
    from apde_mvs_b200.scene import make_plane_scene
    b.build_host()
    d = tmp_path_factory.mktemp("scan1")
    scene = make_plane_scene(256, 192, num_views=5, num_src=4, seed=1)
    scene.write_dense_folder(str(d))
    return d, scene


def test_apd_cli_end_to_end(dense):
    """reference CLI contract (main.cpp:9-24, 180-190, APD.cpp:1225): maps per view + fused APD.ply, through run.py"""
    from apde_mvs_b200.scene import depth_accuracy
    d, scene = dense
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "run.py"), "--data_dir", str(d.parent), "--scans", d.name])
    log = open(d / "APD" / "log.txt").read()
    assert rc == 0, log[-2000:]
    assert "Round nums: 1" in log and "All done" in log and "RunPatchMatch time" in log
    for v in range(5):
        depth = _read_bin(d / "APD" / ("%08d" % v) / "depths.bin")
        normal = _read_bin(d / "APD" / ("%08d" % v) / "normals.bin")
        weak = _read_bin(d / "APD" / ("%08d" % v) / "weak.bin")
        conf = _read_bin(d / "APD" / ("%08d" % v) / "confidence.bin")
        assert depth.shape == (192, 256) and normal.shape == (192, 256, 3) and weak.shape == conf.shape == (192, 256)
        acc = depth_accuracy(depth[12:-12, 12:-12], scene.gt_depth[v][12:-12, 12:-12])
        assert acc >= 0.99, (v, acc)
        assert conf.max() > 1  # geometric passes ran
        import cv2
        skip = cv2.imread(str(d / "APD" / ("%08d" % v) / "skip.png"), cv2.IMREAD_UNCHANGED)  # WeakVisFilter mask, APD.cpp:1035
        assert skip is not None and skip.shape == (192, 256) and set(np.unique(skip)) <= {0, 255}
    raw = open(d / "APD" / "APD.ply", "rb").read()
    head, body = raw.split(b"end_header\n")
    n = int(head.split(b"element vertex ")[1].split(b"\n")[0])
    assert n > 20000 and len(body) == n * 15  # float xyz + uchar bgr
    xyz = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))["p"]
    res = np.abs(4 + 0.15 * xyz[:, 0] - 0.1 * xyz[:, 1] - xyz[:, 2])
    assert np.quantile(res, 0.99) < 0.02  # fused points lie on the synthetic plane z = 4 + 0.15x - 0.1y
    # --only_fuse true re-reads the .bin maps and reproduces the same cloud
    apd = os.path.join(ROOT, "apde_mvs_b200", "_build", "apd")
    os.rename(d / "APD" / "APD.ply", d / "APD" / "first.ply")
    subprocess.check_call([apd, "-d", str(d), "--only_fuse", "true"], stdout=subprocess.DEVNULL)
    assert open(d / "APD" / "APD.ply", "rb").read() == open(d / "APD" / "first.ply", "rb").read()
    # bad usage: missing required option -> usage + non-zero exit (main.cpp:35-39)
    assert subprocess.call([apd], stdout=subprocess.DEVNULL) != 0


def test_apd_class_facade(dense):
    """the reference's ProcessProblem call sequence against the C++ APD class (APD.h:88-114)"""
    d, _ = dense
    out = subprocess.check_output([os.path.join(ROOT, "apde_mvs_b200", "_build", "test_apd_class"), str(d)], text=True)
    assert out.strip().endswith("OK"), out[-1500:]


def test_apd_cli_jpeg_inputs_and_tat_fusion(dense, tmp_path):
    """JPEG images (the reference datasets' format, decoded as cv::imread does) and the Tanks-and-Temples fusion variants
    selected by --dataset (main.cpp:277-283, 294-298)"""
    import cv2
    import shutil
    from apde_mvs_b200.scene import depth_accuracy
    d, scene = dense
    apd = os.path.join(ROOT, "apde_mvs_b200", "_build", "apd")
    counts = {}
    for dataset in ("TaT_i", "TaT_a"):
        j = tmp_path / ("jpg_" + dataset)
        os.makedirs(j / "images")
        shutil.copytree(d / "cams", j / "cams")
        shutil.copy(d / "pair.txt", j / "pair.txt")
        for v in range(5):
            cv2.imwrite(str(j / "images" / ("%08d.jpg" % v)), scene.images[v], [cv2.IMWRITE_JPEG_QUALITY, 95])
        out = subprocess.run([apd, "-d", str(j), "--dataset", dataset], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:]
        depth = _read_bin(j / "APD" / "00000002" / "depths.bin")
        acc = depth_accuracy(depth[12:-12, 12:-12], scene.gt_depth[2][12:-12, 12:-12])
        assert acc >= 0.98, acc
        raw = open(j / "APD" / "APD.ply", "rb").read()
        head, body = raw.split(b"end_header\n")
        n = int(head.split(b"element vertex ")[1].split(b"\n")[0])
        xyz = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))["p"]
        res = np.abs(4 + 0.15 * xyz[:, 0] - 0.1 * xyz[:, 1] - xyz[:, 2])
        print("%s: %d points, plane residual q99 %.4f, view 2 accuracy %.4f" % (dataset, n, np.quantile(res, 0.99), acc))
        assert n > 20000 and np.quantile(res, 0.99) < 0.05
        counts[dataset] = n
    assert counts["TaT_a"] != counts["TaT_i"]  # different thresholds, different clouds
