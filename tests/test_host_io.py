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
