"""CPU tests of the C++ host I/O layer (csrc/host/apd_io.*): the reference's file formats without OpenCV / Boost."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def tool():
    from apde_mvs_b200 import build as b
    b.build_host()
    return os.path.join(ROOT, "apde_mvs_b200", "_build", "test_io")


def _read_pgm(p):
    with open(p, "rb") as f:
        assert f.readline().strip() == b"P5"
        w, h = map(int, f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(), np.uint8).reshape(h, w)


def _read_bin(p):
    with open(p, "rb") as f:
        version, rows, cols, typ = struct.unpack("<4i", f.read(16))
        assert version == 1
        dt = {0: (np.uint8, 1), 16: (np.uint8, 3), 4: (np.int32, 1), 5: (np.float32, 1), 21: (np.float32, 3)}[typ]
        a = np.frombuffer(f.read(), dt[0])
        return a.reshape(rows, cols, dt[1]) if dt[1] > 1 else a.reshape(rows, cols), typ


def test_png_decode_matches_opencv(tool, tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    gray = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    color = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "g.png"), gray)
    cv2.imwrite(str(tmp_path / "c.png"), color)
    subprocess.check_call([tool, "gray", str(tmp_path / "g.png"), str(tmp_path / "g.pgm")])
    assert np.array_equal(_read_pgm(tmp_path / "g.pgm"), gray)
    # colour file read as grey == cv::imread(IMREAD_GRAYSCALE) (APD.cpp:145)
    subprocess.check_call([tool, "gray", str(tmp_path / "c.png"), str(tmp_path / "cg.pgm")])
    assert np.array_equal(_read_pgm(tmp_path / "cg.pgm"), cv2.imread(str(tmp_path / "c.png"), cv2.IMREAD_GRAYSCALE))
    subprocess.check_call([tool, "color", str(tmp_path / "c.png"), str(tmp_path / "c.bin")])
    got, typ = _read_bin(tmp_path / "c.bin")
    assert typ == 16 and np.array_equal(got, cv2.imread(str(tmp_path / "c.png"), cv2.IMREAD_COLOR))
    # WritePNG (skip.png, APD.cpp:1035) is readable by OpenCV and lossless
    subprocess.check_call([tool, "png", str(tmp_path / "c.png"), str(tmp_path / "c2.png")])
    assert np.array_equal(cv2.imread(str(tmp_path / "c2.png"), cv2.IMREAD_COLOR), color)
    subprocess.check_call([tool, "graypng", str(tmp_path / "g.png"), str(tmp_path / "g2.png")])
    assert np.array_equal(cv2.imread(str(tmp_path / "g2.png"), cv2.IMREAD_UNCHANGED), gray)


def _natural(w, h, rng):
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 80 * np.sin(x / 17.0 + c) * np.cos(y / 23.0 - c) + 30 * np.sin((x + y) / 5.0) for c in range(3)], -1)
    return np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)


def test_jpeg_decode_matches_opencv(tool, tmp_path):
    """ReadImage / ReadImageColor on JPEG == cv::imread (libjpeg-turbo defaults: islow IDCT, luma plane for
    IMREAD_GRAYSCALE, fancy chroma up-sampling, fixed-point YCbCr->RGB), bit for bit (APD.cpp:145, 1092)"""
    import cv2
    rng = np.random.default_rng(0)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    variants = {"q90": [cv2.IMWRITE_JPEG_QUALITY, 90], "q35": [cv2.IMWRITE_JPEG_QUALITY, 35], "progressive": [cv2.IMWRITE_JPEG_PROGRESSIVE, 1],
                "restart": [cv2.IMWRITE_JPEG_RST_INTERVAL, 3], "optimized": [cv2.IMWRITE_JPEG_OPTIMIZE, 1],
                "444": [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444], "422": [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422],
                "420": [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420], "440": [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440], "grey": None}
    for (w, h) in ((64, 48), (131, 97), (17, 9), (5, 3)):
        for name, params in variants.items():
            img = _natural(w, h, rng)
            p = str(tmp_path / ("%dx%d_%s.jpg" % (w, h, name)))
            if params is None:
                cv2.imwrite(p, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
            else:
                cv2.imwrite(p, img, params)
            subprocess.check_call([tool, "gray", p, str(tmp_path / "g.pgm")])
            subprocess.check_call([tool, "color", p, str(tmp_path / "c.bin")])
            assert np.array_equal(_read_pgm(tmp_path / "g.pgm"), cv2.imread(p, cv2.IMREAD_GRAYSCALE)), (w, h, name, "grey")
            assert np.array_equal(_read_bin(tmp_path / "c.bin")[0], cv2.imread(p, cv2.IMREAD_COLOR)), (w, h, name, "colour")
    # EXIF orientation: cv::imread turns the image upright
    cv2.imwrite(str(tmp_path / "base.jpg"), _natural(96, 64, rng), [cv2.IMWRITE_JPEG_QUALITY, 92])
    raw = open(tmp_path / "base.jpg", "rb").read()
    for o in range(1, 9):
        for e in "<>":
            tiff = (b"II*\x00" if e == "<" else b"MM\x00*") + struct.pack(e + "I", 8) + struct.pack(e + "H", 1) + \
                struct.pack(e + "HHIHH", 0x0112, 3, 1, o, 0) + struct.pack(e + "I", 0)
            app1 = b"Exif\x00\x00" + tiff
            p = str(tmp_path / ("exif%d.jpg" % o))
            open(p, "wb").write(raw[:2] + b"\xff\xe1" + struct.pack(">H", len(app1) + 2) + app1 + raw[2:])
            subprocess.check_call([tool, "gray", p, str(tmp_path / "g.pgm")])
            subprocess.check_call([tool, "color", p, str(tmp_path / "c.bin")])
            assert np.array_equal(_read_pgm(tmp_path / "g.pgm"), cv2.imread(p, cv2.IMREAD_GRAYSCALE)), ("exif", o, e)
            assert np.array_equal(_read_bin(tmp_path / "c.bin")[0], cv2.imread(p, cv2.IMREAD_COLOR)), ("exif", o, e)
    # truncated / non-JPEG data is an error, not a crash
    open(tmp_path / "bad.jpg", "wb").write(raw[:200])
    assert subprocess.run([tool, "gray", str(tmp_path / "bad.jpg"), str(tmp_path / "g.pgm")], capture_output=True).returncode in (0, 1)
    open(tmp_path / "bad2.jpg", "wb").write(b"not an image")
    assert subprocess.run([tool, "gray", str(tmp_path / "bad2.jpg"), str(tmp_path / "g.pgm")], capture_output=True).returncode == 1


def test_bin_mat_format_roundtrip(tool, tmp_path):
    """the 'dmb' layout of ReadBinMat / WriteBinMat (APD.cpp:18-83): int32 version=1, rows, cols, cv type, raw rows"""
    rng = np.random.default_rng(1)
    for typ, arr in ((5, rng.random((11, 7), dtype=np.float32)), (21, rng.random((5, 9, 3), dtype=np.float32)),
                     (0, rng.integers(0, 3, (8, 8), dtype=np.uint8)), (4, rng.integers(-5, 5, (3, 4)).astype(np.int32))):
        src = tmp_path / ("in%d.bin" % typ)
        with open(src, "wb") as f:
            f.write(struct.pack("<4i", 1, arr.shape[0], arr.shape[1], typ))
            f.write(arr.tobytes())
        out = subprocess.check_output([tool, "bin", str(src), str(tmp_path / "out.bin")], text=True)
        assert out.split() == [str(arr.shape[0]), str(arr.shape[1]), str(typ)]
        got, t2 = _read_bin(tmp_path / "out.bin")
        assert t2 == typ and np.array_equal(got, arr)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(struct.pack("<4i", 2, 1, 1, 5) + b"\0\0\0\0")
    assert subprocess.call([tool, "bin", str(bad), str(tmp_path / "o.bin")], stdout=subprocess.DEVNULL) != 0  # version check


def test_camera_pairs_ply(tool, tmp_path):
    from apde_mvs_b200.scene import make_plane_scene
    scene = make_plane_scene(64, 48, num_views=4, num_src=3, seed=2)
    scene.write_dense_folder(str(tmp_path))
    vals = list(map(float, subprocess.check_output([tool, "cam", str(tmp_path / "cams" / "00000002_cam.txt")], text=True).split()))
    cam = scene.cameras[2]
    want = list(cam.K) + list(cam.R) + list(cam.t) + list(cam.c) + [cam.depth_min, cam.depth_max, cam.interval, cam.depth_num]
    assert np.allclose(vals, want, rtol=2e-6, atol=1e-6)
    # default depth_num / depth_max when the last line only has two numbers (APD.cpp:121-124)
    txt = open(tmp_path / "cams" / "00000002_cam.txt").read().strip().split("\n")
    txt[-1] = "2.5 0.01"
    (tmp_path / "short_cam.txt").write_text("\n".join(txt) + "\n")
    v2 = list(map(float, subprocess.check_output([tool, "cam", str(tmp_path / "short_cam.txt")], text=True).split()))
    assert v2[-1] == 192 and abs(v2[-3] - (0.01 * 192 + 2.5)) < 1e-5
    # pair.txt: score <= 0 entries are dropped (main.cpp:80-82)
    lines = open(tmp_path / "pair.txt").read().split("\n")
    lines[2] = "3 1 1.0 2 0.0 3 0.5"
    (tmp_path / "pair.txt").write_text("\n".join(lines))
    out = subprocess.check_output([tool, "pairs", str(tmp_path)], text=True).strip().split("\n")
    assert out[0] == "0 .png: 1 3" and len(out) == 4
    assert os.path.isdir(tmp_path / "APD" / "00000003")
    subprocess.check_call([tool, "ply", str(tmp_path / "p.ply")])
    raw = open(tmp_path / "p.ply", "rb").read()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 2" in head and b"property uchar blue" in head and b"format binary_little_endian 1.0" in head
    assert len(body) == 2 * 15 and struct.unpack("<3f3B", body[:15]) == (1.5, -2.0, 3.25, 10, 20, 30)


def _write_bin(p, a, typ):
    with open(p, "wb") as f:
        f.write(struct.pack("<4i", 1, a.shape[0], a.shape[1], typ))
        f.write(np.ascontiguousarray(a).tobytes())


def test_jpeg_writer_is_decodable_and_close(tool, tmp_path):
    """WriteJPEG (depth_k.jpg / normal_k.jpg, APD.cpp:228,286): baseline 4:4:4 files that libjpeg decodes, as close to the
    source as cv2's own quality-95 files"""
    import cv2
    rng = np.random.default_rng(5)
    for name, (w, h) in {"a": (67, 45), "b": (64, 64), "c": (1, 1), "d": (17, 8)}.items():
        img = _natural(w, h, rng)
        cv2.imwrite(str(tmp_path / f"{name}.png"), img)
        subprocess.check_call([tool, "jpg", str(tmp_path / f"{name}.png"), str(tmp_path / f"{name}.jpg")])
        got = cv2.imread(str(tmp_path / f"{name}.jpg"), cv2.IMREAD_COLOR)
        assert got.shape == img.shape
        cv2.imwrite(str(tmp_path / f"{name}_cv.jpg"), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        ref = cv2.imread(str(tmp_path / f"{name}_cv.jpg"), cv2.IMREAD_COLOR)
        err, err_cv = np.abs(got.astype(int) - img).mean(), np.abs(ref.astype(int) - img).mean()
        assert err <= err_cv + (0.25 if w * h > 64 else 1.0) and np.abs(got.astype(int) - img).max() <= np.abs(ref.astype(int) - img).max() + 4, (name, err, err_cv)
        # our own decoder reads it back identically to libjpeg
        subprocess.check_call([tool, "color", str(tmp_path / f"{name}.jpg"), str(tmp_path / f"{name}.bin")])
        assert np.array_equal(_read_bin(tmp_path / f"{name}.bin")[0], got)
        gray = img[..., 1].copy()
        cv2.imwrite(str(tmp_path / f"{name}_g.png"), gray)
        subprocess.check_call([tool, "grayjpg", str(tmp_path / f"{name}_g.png"), str(tmp_path / f"{name}_g.jpg")])
        gg = cv2.imread(str(tmp_path / f"{name}_g.jpg"), cv2.IMREAD_UNCHANGED)
        assert gg.shape == gray.shape and np.abs(gg.astype(int) - gray).max() <= 10
    # hard case: white noise at quality 100 (large coefficients, long zero runs absent)
    noise = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "n.png"), noise)
    subprocess.check_call([tool, "jpg", str(tmp_path / "n.png"), str(tmp_path / "n.jpg"), "100"])
    assert np.abs(cv2.imread(str(tmp_path / "n.jpg")).astype(int) - noise).max() <= 6
    # flat picture: every AC coefficient zero (EOB-only blocks) and 16+ zero runs
    flat = np.full((24, 40, 3), 200, np.uint8)
    flat[20:, 36:] = 10
    cv2.imwrite(str(tmp_path / "f.png"), flat)
    subprocess.check_call([tool, "jpg", str(tmp_path / "f.png"), str(tmp_path / "f.jpg")])
    assert np.abs(cv2.imread(str(tmp_path / "f.jpg")).astype(int) - flat)[:16, :32].max() <= 1


def test_show_pictures_follow_the_reference(tool, tmp_path):
    """ShowDepthMap / ShowNormalMap / ShowWeakImage / ShowConfidenceMap (APD.cpp:162-314) against OpenCV's own operators"""
    import cv2
    rng = np.random.default_rng(9)
    h, w = 48, 72
    y, x = np.mgrid[0:h, 0:w]
    depth = (4 + 0.02 * x - 0.015 * y + 0.3 * np.sin(x / 6.0)).astype(np.float32)
    depth[rng.random((h, w)) < 0.1] = 0  # invalid pixels are excluded from the statistics
    normal = rng.normal(size=(h, w, 3)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=2, keepdims=True)
    normal[depth == 0] = 0
    weak = rng.integers(0, 3, (h, w), dtype=np.uint8)
    conf = rng.integers(3, 200, (h, w), dtype=np.uint8)
    _write_bin(tmp_path / "depths.bin", depth, 5)
    _write_bin(tmp_path / "normals.bin", normal, 21)
    _write_bin(tmp_path / "weak.bin", weak, 0)
    _write_bin(tmp_path / "confidence.bin", conf, 0)
    dmin, dmax = 2.0 * 0.6, 8.0 * 1.2
    subprocess.check_call([tool, "show", str(tmp_path), repr(dmin), repr(dmax)])
    # the colour map is cv::COLORMAP_JET entry for entry
    jet, _ = _read_bin(tmp_path / "jet.bin")
    assert np.array_equal(jet.reshape(256, 3), cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET).reshape(256, 3))
    # depth: mean of column means, sigma from the mean of column variances, ramp over mean +- 2 sigma, jet (APD.cpp:164-227)
    ok = (depth >= dmin) & (depth <= dmax)
    cols = [c for c in range(w) if ok[:, c].any()]
    mean = np.float32(np.mean([depth[ok[:, c], c].mean() for c in cols]))
    sigma = np.sqrt(np.mean([((depth[ok[:, c], c] - mean) ** 2).mean() for c in cols]))
    lo = mean - 2 * sigma
    gray = (np.clip((depth - lo) / (4 * sigma), 0, 1) * 255).astype(np.uint8)
    want = cv2.applyColorMap(gray, cv2.COLORMAP_JET)
    got = cv2.imread(str(tmp_path / "depth_3.jpg"), cv2.IMREAD_COLOR)
    cv2.imwrite(str(tmp_path / "depth_cv.jpg"), want)
    base = np.abs(cv2.imread(str(tmp_path / "depth_cv.jpg")).astype(int) - want).mean()
    assert got.shape == want.shape and np.abs(got.astype(int) - want).mean() <= base + 0.5
    # normals: unit vector * 127.5 + 127.5, rounded as Mat::convertTo does
    want_n = np.empty((h, w, 3), np.uint8)
    cv2.convertScaleAbs(normal, want_n, 1)  # placeholder to get the array type; real values below
    want_n = np.clip(np.rint(normal * np.float32(127.5) + np.float32(127.5)), 0, 255).astype(np.uint8)
    got_n = cv2.imread(str(tmp_path / "normal_3.jpg"), cv2.IMREAD_COLOR)
    cv2.imwrite(str(tmp_path / "normal_cv.jpg"), want_n)
    base_n = np.abs(cv2.imread(str(tmp_path / "normal_cv.jpg")).astype(int) - want_n).mean()
    assert np.abs(got_n.astype(int) - want_n).mean() <= base_n + 0.5
    # the two PNGs are lossless: exact
    colours = np.array([[255, 255, 255], [0, 255, 0], [0, 0, 255]], np.uint8)  # WEAK, STRONG, UNKNOWN in BGR (APD.cpp:298-308)
    assert np.array_equal(cv2.imread(str(tmp_path / "weak_3.png"), cv2.IMREAD_COLOR), colours[weak])
    lo_c, hi_c = int(conf.min()), int(conf.max())
    want_c = ((conf.astype(int) - lo_c) * 255 // max(hi_c - lo_c, 1)).astype(np.uint8)
    assert np.array_equal(cv2.imread(str(tmp_path / "confidence_3.png"), cv2.IMREAD_UNCHANGED), want_c)


def test_threaded_view_loader_matches_serial_decoding(tool, tmp_path):
    """LoadViews (the scene loader of the `apd` CLI: one decoding thread per view, batches of min(cores, views, 16)) returns,
    in view order, exactly what ReadImage / ReadImageColor / ReadCamera / ReadBinMat give one by one -- 19 views, so that more
    than one batch runs; JPEG and PNG inputs, a label map for some views only, one undecodable image"""
    import cv2
    rng = np.random.default_rng(11)
    d = tmp_path / "scan"
    os.makedirs(d / "images"); os.makedirs(d / "cams"); os.makedirs(d / "sa_masks")
    V, w, h = 19, 96, 64
    imgs, labels = [], {}
    for v in range(V):
        img = _natural(w, h, rng)
        ext = ".jpg" if v % 2 else ".png"
        if v != 7:
            cv2.imwrite(str(d / "images" / ("%08d%s" % (v, ext))), img)
        else:
            open(d / "images" / ("%08d%s" % (v, ext)), "wb").write(b"not an image at all")
        imgs.append(str(d / "images" / ("%08d%s" % (v, ext))))
        with open(d / "cams" / ("%08d_cam.txt" % v), "w") as f:
            f.write("extrinsic\n1 0 0 %d\n0 1 0 0\n0 0 1 0\n0 0 0 1\n\nintrinsic\n%d 0 48\n0 %d 32\n0 0 1\n\n%g 0.01 192 3.5\n" % (v, 100 + v, 100 + v, 1.5 + v))
        if v % 3 == 0:
            labels[v] = rng.integers(0, 9, (h // 2, w // 2), dtype=np.uint8)
            _write_bin(d / "sa_masks" / ("%08d.bin" % v), labels[v], 0)
    with open(d / "pair.txt", "w") as f:
        f.write("%d\n" % V)
        for v in range(V):
            f.write("%d\n2 %d 1.0 %d 1.0\n" % (v, (v + 1) % V, (v + 2) % V))
    # view 7 is not decodable: the loader reports image_ok = 0 for it (the CLI then stops with the reference's message)
    out = subprocess.run([tool, "load", str(d)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l and l[0].isdigit() and len(l.split()) == 9]
    assert [int(l[0]) for l in lines] == list(range(V))  # view order kept across batches

    def chk(a):
        s = 0
        for b in np.ascontiguousarray(a).ravel().tolist():
            s = (s * 1315423911 + b) & 0xFFFFFFFFFFFFFFFF
        return s
    for v, l in enumerate(lines):
        if v == 7:
            assert l[1] == "0"
            continue
        assert l[1] == "1" and l[2] == "1" and l[3] == "%dx%d" % (w, h)
        assert int(l[4]) == chk(cv2.imread(imgs[v], cv2.IMREAD_GRAYSCALE))
        assert int(l[5]) == chk(cv2.imread(imgs[v], cv2.IMREAD_COLOR))
        assert int(l[6]) == (chk(labels[v]) if v in labels else 0)
        assert float(l[7]) == 100 + v and abs(float(l[8]) - (1.5 + v)) < 1e-6


def test_show_writer_threads_write_the_same_pictures(tool, tmp_path):
    """ShowWriter (the CLI's asynchronous Show*: worker threads encode while the next round runs): 20 views, more than the
    thread limit, every view's four files complete when the writer is destroyed and identical to the one-by-one result"""
    rng = np.random.default_rng(3)
    h, w = 40, 56
    y, x = np.mgrid[0:h, 0:w]
    depth = (4 + 0.02 * x - 0.015 * y).astype(np.float32)
    normal = rng.normal(size=(h, w, 3)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=2, keepdims=True)
    _write_bin(tmp_path / "depths.bin", depth, 5)
    _write_bin(tmp_path / "normals.bin", normal, 21)
    _write_bin(tmp_path / "weak.bin", rng.integers(0, 3, (h, w), dtype=np.uint8), 0)
    _write_bin(tmp_path / "confidence.bin", rng.integers(3, 200, (h, w), dtype=np.uint8), 0)
    n = 20
    subprocess.check_call([tool, "showasync", str(tmp_path), str(n), "1.2", "9.6"])
    first = None
    for k in range(n):
        d = tmp_path / str(k)
        for name in ("depth_3.jpg", "normal_3.jpg", "weak_3.png", "confidence_3.png"):
            assert (d / name).exists() and os.path.getsize(d / name) > 100, (k, name)
        assert open(d / "depth_3.jpg", "rb").read() == open(d / "depth_sync.jpg", "rb").read()
        first = first or open(d / "normal_3.jpg", "rb").read()
        assert open(d / "normal_3.jpg", "rb").read() == first
