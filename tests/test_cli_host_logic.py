"""CPU test of the `apd` CLI's HOST logic, end to end, against a test double of libapde.so (tests/mock/mock_apde.c).

The double only check-sums uploads and returns synthetic maps, so everything that IS real here is host code of the drop-in
surface: pair.txt / camera / image / label-map loading on threads and upload in view order (APD.cpp: SceneSession), the pass loop
of main.cpp:303-367, Show* pictures through the worker threads, the four .bin maps, skip.png, the fused .ply, --only_fuse,
--no_fuse, the TaT switches.  The product binary is not involved (its libapde.so needs a CUDA device and has no CPU path)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import ROOT

HOST = os.path.join(ROOT, "apde_mvs_b200", "csrc", "host")


def _chk(a):
    s = 0
    for b in np.ascontiguousarray(a).ravel().tolist():
        s = (s * 1315423911 + b) & 0xFFFFFFFFFFFFFFFF
    return s


def _read_bin(p):
    with open(p, "rb") as f:
        version, rows, cols, typ = struct.unpack("<4i", f.read(16))
        dt, ch = {0: (np.uint8, 1), 5: (np.float32, 1), 21: (np.float32, 3)}[typ]
        a = np.frombuffer(f.read(), dt)
        return a.reshape(rows, cols, ch) if ch > 1 else a.reshape(rows, cols)


@pytest.fixture(scope="module")
def mock_apd(tmp_path_factory):
    d = tmp_path_factory.mktemp("mockbin")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", str(d / "libapde.so"),
                           os.path.join(ROOT, "tests", "mock", "mock_apde.c")])
    srcs = [os.path.join(HOST, f) for f in ("apd_io.cpp", "apd_jpeg.cpp", "apd_show.cpp", "APD.cpp", "main.cpp")]
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", str(d / "apd")] + srcs + ["-L" + str(d), "-lapde", "-lz", "-Wl,-rpath,$ORIGIN"])
    return str(d / "apd")


def _make_folder(d, V, w, h, rng, labels_for):
    import cv2
    os.makedirs(d / "images"); os.makedirs(d / "cams")
    if labels_for:
        os.makedirs(d / "sa_masks")
    imgs, labels = [], {}
    y, x = np.mgrid[0:h, 0:w]
    for v in range(V):
        img = np.stack([128 + 80 * np.sin(x / 17.0 + c + v) * np.cos(y / 23.0 - c) + 30 * np.sin((x + y) / 5.0) for c in range(3)], -1).clip(0, 255).astype(np.uint8)
        p = str(d / "images" / ("%08d%s" % (v, ".jpg" if v % 2 else ".png")))
        cv2.imwrite(p, img)
        imgs.append(p)
        with open(d / "cams" / ("%08d_cam.txt" % v), "w") as f:
            f.write("extrinsic\n1 0 0 %d\n0 1 0 0\n0 0 1 0\n0 0 0 1\n\nintrinsic\n%d 0 %d\n0 %d %d\n0 0 1\n\n%g 0.01 192 9.5\n" % (v, 100 + v, w // 2, 100 + v, h // 2, 1.5 + v))
        if v in labels_for:
            labels[v] = rng.integers(0, 9, (h // 2, w // 2), dtype=np.uint8)
            with open(d / "sa_masks" / ("%08d.bin" % v), "wb") as f:
                f.write(struct.pack("<4i", 1, h // 2, w // 2, 0))
                f.write(labels[v].tobytes())
    with open(d / "pair.txt", "w") as f:
        f.write("%d\n" % V)
        for v in range(V):
            f.write("%d\n3 %d 1.0 %d 0.5 %d -1.0\n" % (v, (v + 1) % V, (v + 2) % V, (v + 3) % V))  # the score <= 0 entry is dropped
    return imgs, labels


def _mock_maps(v, w, h, passes):
    y, x = np.mgrid[0:h, 0:w]
    depth = (np.float32(2.0) + np.float32(0.01) * x.astype(np.float32) + np.float32(0.02) * y.astype(np.float32) + np.float32(0.1) * np.float32(v)
             + np.float32(0.001) * np.float32(passes)).astype(np.float32)
    return depth, ((x + y + v) % 3).astype(np.uint8), ((x * 3 + y + v) % 200).astype(np.uint8)


def test_cli_host_flow_against_the_test_double(mock_apd, tmp_path):
    import cv2
    rng = np.random.default_rng(4)
    V, w, h = 19, 96, 64  # 19 views: more than one loader batch on a 16-core box
    d = tmp_path / "scan"
    imgs, labels = _make_folder(d, V, w, h, rng, labels_for={0, 3, 6, 18})
    out = subprocess.run([mock_apd, "-d", str(d), "--dataset", "General"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    log = out.stdout
    # ---- loading: every view uploaded once, in view order, with the pixels cv::imread yields; label maps where they exist
    views = [l.split() for l in log.splitlines() if l.startswith("MOCK view")]
    assert [int(l[2]) for l in views] == list(range(V))
    for v, l in enumerate(views):
        assert int(l[4]) == _chk(cv2.imread(imgs[v], cv2.IMREAD_GRAYSCALE)) and int(l[6]) == _chk(cv2.imread(imgs[v], cv2.IMREAD_COLOR))
        assert float(l[8]) == 100 + v and abs(float(l[10]) - (1.5 + v)) < 1e-6
    sa = {int(l.split()[2]): l.split() for l in log.splitlines() if l.startswith("MOCK sa")}
    assert sorted(sa) == sorted(labels) and all(sa[v][3] == "%dx%d" % (w // 2, h // 2) and int(sa[v][4]) == _chk(labels[v]) for v in labels)
    assert "sa masks: 4 of %d views" % V in log
    pairs = [l for l in log.splitlines() if l.startswith("MOCK pairs")]
    assert pairs[0] == "MOCK pairs 0: 1 2" and pairs[V - 1] == "MOCK pairs %d: 0 1" % (V - 1)  # score -1 dropped (main.cpp:76-84)
    assert log.index("MOCK commit") > log.index("MOCK pairs %d:" % (V - 1))
    # ---- the pass loop: one round of 1 + 3 passes at this size, reference banner lines
    assert [l for l in log.splitlines() if l.startswith("MOCK pass")] == ["MOCK pass %d use_sa 1 geom_factor 0.2" % p for p in range(4)]
    assert "Round nums: 1" in log and "======== iteration 3========" in log and "All done" in log
    # ---- outputs: maps, pictures of iteration 3 (none for the other iterations), skip.png, fused cloud
    colours = np.array([[255, 255, 255], [0, 255, 0], [0, 0, 255]], np.uint8)
    for v in range(V):
        r = d / "APD" / ("%08d" % v)
        depth, weak, conf = _mock_maps(v, w, h, 4)
        assert np.array_equal(_read_bin(r / "depths.bin"), depth) and np.array_equal(_read_bin(r / "weak.bin"), weak)
        assert np.array_equal(_read_bin(r / "confidence.bin"), conf) and _read_bin(r / "normals.bin").shape == (h, w, 3)
        assert np.array_equal(cv2.imread(str(r / "weak_3.png"), cv2.IMREAD_COLOR), colours[weak])
        assert np.array_equal(cv2.imread(str(r / "confidence_3.png"), cv2.IMREAD_UNCHANGED), ((conf.astype(int) - conf.min()) * 255 // int(conf.max() - conf.min())).astype(np.uint8))
        assert cv2.imread(str(r / "depth_3.jpg")).shape == (h, w, 3) and cv2.imread(str(r / "normal_3.jpg")).shape == (h, w, 3)
        assert not (r / "depth_0.jpg").exists() and not (r / "depth_2.jpg").exists()
        skip = cv2.imread(str(r / "skip.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(skip.ravel(), np.where((np.arange(w * h) + v) % 7 == 0, 255, 0).astype(np.uint8))
    raw = open(d / "APD" / "APD.ply", "rb").read()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 5" in head and len(body) == 5 * 15
    pts = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))
    assert np.allclose(pts["p"][2], [2.0, 2.25, 2.5]) and list(pts["c"][2]) == [22, 23, 24]  # use_weak_filter = KEEP (2) after the filter ran
    assert log.count("MOCK fusion run") == 1  # count + take: the cloud is fused once, not once to count and once to fill
    # ---- --only_fuse re-reads the maps; TaT switches: geom_factor 0.05 and the other fusion variants
    out2 = subprocess.run([mock_apd, "-d", str(d), "--only_fuse", "true", "--dataset", "TaT_a"], capture_output=True, text=True)
    assert out2.returncode == 0 and "MOCK upload 18 %dx%d d00 %.6g" % (w, h, _mock_maps(18, w, h, 4)[0][0, 0]) in out2.stdout
    assert "MOCK pass" not in out2.stdout and b"element vertex 7" in open(d / "APD" / "APD.ply", "rb").read(300)
    assert out2.stdout.count("MOCK fusion run variant 2") == 1
    for f in ("depth_3.jpg", "weak_3.png"):
        os.remove(d / "APD" / "00000000" / f)
    os.remove(d / "APD" / "APD.ply")
    out3 = subprocess.run([mock_apd, "-d", str(d), "--no_fuse", "true", "--dataset", "TaT_i", "--use_sa", "false"], capture_output=True, text=True,
                          env=dict(os.environ, APDE_NO_SHOW="1"))
    assert out3.returncode == 0 and "MOCK pass 3 use_sa 0 geom_factor 0.05" in out3.stdout and "Skip fusion, all done!" in out3.stdout
    assert not (d / "APD" / "APD.ply").exists() and not (d / "APD" / "00000000" / "depth_3.jpg").exists()
    # ---- --only_fuse validates the maps it is about to upload (APD.cpp:1106-1119 checks sizes; types and truncation as well here)
    import struct
    victim = d / "APD" / "00000003" / "weak.bin"
    good = open(victim, "rb").read()
    for bad_bytes, msg in ((struct.pack("<4i", 1, h, w - 1, 0) + bytes(h * (w - 1)), "weak size is not equal to depth size"),
                           (struct.pack("<4i", 1, h, w, 5) + bytes(4 * h * w), "unexpected element type"),
                           (struct.pack("<4i", 1, 1 << 15, 1 << 15, 0) + bytes(16), "is truncated"),
                           (struct.pack("<4i", 1, -4, w, 0), "Size error")):
        open(victim, "wb").write(bad_bytes)
        bad = subprocess.run([mock_apd, "-d", str(d), "--only_fuse", "true"], capture_output=True, text=True)
        assert bad.returncode != 0 and msg in bad.stdout and "MOCK fusion run" not in bad.stdout, (msg, bad.stdout[-400:])
    open(victim, "wb").write(good)
    # ---- errors: an undecodable image stops the run with the reference's message (main.cpp:104-127)
    open(imgs[5], "wb").write(b"garbage")
    out4 = subprocess.run([mock_apd, "-d", str(d)], capture_output=True, text=True)
    assert out4.returncode != 0 and "Images may error, check it!" in out4.stdout


def test_cli_two_rounds_write_pictures_per_round(mock_apd, tmp_path):
    """a size above 800 px gives two rounds (ComputeRoundNum): pictures for iterations 3 and 7 only, maps written once at the end"""
    rng = np.random.default_rng(5)
    d = tmp_path / "scan"
    _make_folder(d, 3, 832, 48, rng, labels_for=set())
    out = subprocess.run([mock_apd, "-d", str(d), "--no_fuse", "true"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-1500:]
    assert "Round nums: 2" in out.stdout and "Can't find sa mask folder" in out.stdout
    assert len([l for l in out.stdout.splitlines() if l.startswith("MOCK pass")]) == 8
    for v in range(3):
        r = d / "APD" / ("%08d" % v)
        names = sorted(os.listdir(r))
        assert [n for n in names if n.startswith("depth_")] == ["depth_3.jpg", "depth_7.jpg"]
        assert [n for n in names if n.startswith("confidence_")] == ["confidence_3.png", "confidence_7.png"]
        assert np.array_equal(_read_bin(r / "depths.bin"), _mock_maps(v, 832, 48, 8)[0])


def test_cli_job_over_two_gpus_against_the_test_double(mock_apd, tmp_path):
    """`apd --gpus 2`: one scene, two contexts, two host threads.  Both contexts get the whole scene (decoded once), both join
    the job with the same id, every rank runs all passes as collective calls, each rank writes the maps / skip.png of ITS block
    of views, the maps are gathered before fusion and rank 0 alone writes the cloud."""
    rng = np.random.default_rng(6)
    V, w, h = 7, 64, 48  # 7 views over 2 ranks: blocks of 4 and 3
    d = tmp_path / "scan"
    _make_folder(d, V, w, h, rng, labels_for={1})
    out = subprocess.run([mock_apd, "-d", str(d), "--gpus", "2", "--gpu_index", "3", "--dataset", "TaT_i"], capture_output=True, text=True,
                         env=dict(os.environ, APDE_NO_SHOW="1"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    log = out.stdout.splitlines()
    assert "MOCK create 0 device 3" in log and "MOCK create 1 device 4" in log
    assert len([l for l in log if l.startswith("MOCK view")]) == V and len([l for l in log if l.startswith("MOCK ctx1+ view")]) == V
    assert "MOCK comm_init ctx 0 rank 0 of 2" in log and "MOCK comm_init ctx 1 rank 1 of 2" in log
    for r in (0, 1):
        assert [l for l in log if l.startswith("MOCK rank %d of 2 pass" % r)] == ["MOCK rank %d of 2 pass %d" % (r, p) for p in range(4)]
        ex = [l for l in log if l.startswith("MOCK rank %d exchange" % r)]
        assert ex == ["MOCK rank %d exchange %d" % (r, k) for k in (1, 2, 3, 4)]  # normal, weak, confidence; skip after the filter
        assert "MOCK rank %d mark %dx%d" % (r, w, h) in log
    assert "There are 7 problems needed to be processed! (2 GPUs)" in log and "All done" in log
    for v in range(V):
        r = d / "APD" / ("%08d" % v)
        depth, weak, conf = _mock_maps(v, w, h, 4)
        assert np.array_equal(_read_bin(r / "depths.bin"), depth) and np.array_equal(_read_bin(r / "weak.bin"), weak)
        import cv2
        skip = cv2.imread(str(r / "skip.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(skip.ravel(), np.where((np.arange(w * h) + v) % 7 == 0, 255, 0).astype(np.uint8))
    head = open(d / "APD" / "APD.ply", "rb").read(300)
    assert b"element vertex 6" in head  # RunFusion_TAT_I (variant 1) on rank 0 only
    # the exports that step through a pass view by view are single-GPU
    out2 = subprocess.run([mock_apd, "-d", str(d), "--gpus", "2", "--export_anchor", "true"], capture_output=True, text=True)
    assert out2.returncode != 0 and "without --gpus" in out2.stdout
