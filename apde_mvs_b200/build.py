"""Build recipe for libapde.so (hand-written CUDA for sm_100a behind the C ABI of include/apde.h).

In-tree build: the .so lands in apde_mvs_b200/_build/ and travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libapde.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["apde_api.cu", "apde_comm.cu", "apde_kernels.cu", "apde_apd.cu", "apde_sweep.cu", "apde_prop.cu", "apde_maps.cu", "apde_lists.cu", "apde_fusion.cu", "apde_microbench.cu"]
# the fusion kernels restate HOST code of the reference (IEEE arithmetic, no FMA contraction): no fast-math there
PRECISE = {"apde_fusion.cu": ["-fmad=false"]}
FAST = ["--use_fast_math"]
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build_variant(name, defines):
    """an experimental build of the same ABI with extra -D flags, into ab/<name>/libapde.so (A/B runs: APDE_LIB=...)"""
    out = os.path.join(HERE, "..", "ab", name)
    os.makedirs(out, exist_ok=True)
    procs, objs = [], []
    for s in SOURCES:
        obj = os.path.join(out, s.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([NVCC] + [f for f in FLAGS if f not in ("-Xptxas", "-v")] + PRECISE.get(s, FAST) + ["-D" + d for d in defines] +
                                      ["-c", os.path.join(CSRC, s), "-o", obj], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    for p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(err.decode()[-3000:])
    lib = os.path.join(out, "libapde.so")
    subprocess.check_call([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", lib] + objs + ["-lcudart", "-ldl"])
    return lib


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "apde.h"))
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            log = open(obj + ".log", "w")
            procs.append((s, subprocess.Popen([NVCC] + FLAGS + PRECISE.get(s, FAST) + ["-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT), log))
    failed = False
    for s, p, log in procs:
        rc = p.wait()
        log.close()
        if rc != 0 or verbose:
            sys.stderr.write(open(os.path.join(OUT_DIR, s.replace(".cu", ".o")) + ".log").read())
        failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(LIB):
        subprocess.check_call([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


HOST_DIR = os.path.join(CSRC, "host")
APD_BIN = os.path.join(OUT_DIR, "apd")
APD_TEST_BIN = os.path.join(OUT_DIR, "test_apd_class")


def build_host():
    """C++17 host side of the drop-in surface: the APD class, the I/O layer and the `apd` CLI, linked against libapde.so"""
    lib = build()
    common = [os.path.join(HOST_DIR, f) for f in ("apd_io.cpp", "apd_jpeg.cpp", "apd_show.cpp", "APD.cpp")]
    hdrs = [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR)]
    for out, main_src in ((APD_BIN, "main.cpp"), (APD_TEST_BIN, "test_apd_class.cpp"), (os.path.join(OUT_DIR, "test_io"), "test_io.cpp")):
        srcs = common + [os.path.join(HOST_DIR, main_src)]
        if os.path.exists(out) and all(os.path.getmtime(h) < os.path.getmtime(out) for h in hdrs + [lib]):
            continue
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-o", out] + srcs +
                              ["-L" + OUT_DIR, "-lapde", "-lz", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-lcudart",
                               "-Wl,-rpath,/usr/local/cuda/lib64"])
    return APD_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host())
