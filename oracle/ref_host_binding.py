"""ctypes binding of oracle/_ref/libapd_ref_host.so: the reference's own HOST code (APD.cpp compiled unmodified) --
TEST INFRASTRUCTURE ONLY.  Runs on the CPU; used to pin the oracle's restatement of fusion and file formats."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libapd_ref_host.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_read_camera.argtypes = [C.c_char_p, C.c_void_p]
        L.ref_copy_bin_mat.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_run_fusion.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.ref_generate_sample_list.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        L.ref_compute_round_num.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def generate_sample_list(dense_folder, max_views=4096, max_src=1 << 20):
    """the reference's GenerateSampleList -> (ref ids, list of source-id lists, image extension)"""
    ids = np.zeros(max_views, np.int32)
    offs = np.zeros(max_views + 1, np.int32)
    src = np.zeros(max_src, np.int32)
    ext = C.create_string_buffer(16)
    n = lib().ref_generate_sample_list(str(dense_folder).encode(), max_views, max_src, ids.ctypes.data, offs.ctypes.data, src.ctypes.data, ext)
    if n < 0:
        raise RuntimeError("reference GenerateSampleList: buffer too small")
    return ids[:n].tolist(), [src[offs[i]:offs[i + 1]].tolist() for i in range(n)], ext.value.decode()


def compute_round_num(dense_folder):
    return lib().ref_compute_round_num(str(dense_folder).encode())


def read_camera(path):
    """the reference's ReadCamera -> dict(K, R, t, c, height, width, depth_min, depth_max, interval, depth_num)"""
    out = np.zeros(30, np.float32)
    if lib().ref_read_camera(str(path).encode(), out.ctypes.data) != 0:
        raise RuntimeError("reference ReadCamera failed on %s" % path)
    return {"K": out[0:9].copy(), "R": out[9:18].copy(), "t": out[18:21].copy(), "c": out[21:24].copy(), "height": int(out[24]),
            "width": int(out[25]), "depth_min": float(out[26]), "depth_max": float(out[27]), "interval": float(out[28]),
            "depth_num": float(out[29])}


def copy_bin_mat(src, dst):
    """ReadBinMat(src) then WriteBinMat(dst) by the reference; returns the cv type code"""
    rc = lib().ref_copy_bin_mat(str(src).encode(), str(dst).encode())
    if rc < 0:
        raise RuntimeError("reference ReadBinMat / WriteBinMat failed (%d)" % rc)
    return rc


def run_fusion(dense_folder, pairs, variant=0, weak_filter=True, export_color=True, img_ext=".ppm", name="ref.ply"):
    """the reference's RunFusion / RunFusion_TAT_I / RunFusion_TAT_A on a dense folder; returns (xyz float32 [n,3], bgr uint8 [n,3])"""
    V = len(pairs)
    ids = np.arange(V, dtype=np.int32)
    offs = np.zeros(V + 1, np.int32)
    flat = []
    for i, p in enumerate(pairs):
        flat += list(p)
        offs[i + 1] = len(flat)
    flat = np.ascontiguousarray(flat, np.int32)
    lib().ref_run_fusion(str(dense_folder).encode(), V, ids.ctypes.data, offs.ctypes.data, flat.ctypes.data, img_ext.encode(), variant,
                         int(weak_filter), int(export_color), name.encode())
    raw = open(os.path.join(str(dense_folder), "APD", name), "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    n = int(head.split(b"element vertex ")[1].split(b"\n")[0])
    if export_color:
        rec = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]), count=n)
        return rec["p"].copy(), rec["c"].copy()
    return np.frombuffer(body, "<f4", count=3 * n).reshape(n, 3).copy(), None
