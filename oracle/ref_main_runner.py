"""Run the reference's OWN main() (oracle/_ref/libapd_ref_full.so: APD.cu + APD.cpp + main.cpp compiled unmodified) in
this process -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/ref_main_runner.py <seed> --dense_folder <dir> [reference flags ...]

Images in <dir>/images must hold binary PGM / PPM content (the stub cv::imread decodes by magic, whatever the extension)."""
import ctypes as C
import os
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libapd_ref_full.so")


def available():
    return os.path.exists(LIB)


def main(argv):
    seed = int(argv[0])
    args = [b"APD"] + [a.encode() for a in argv[1:]]
    lib = C.CDLL(LIB)
    lib.ref_full_set_seed.argtypes = [C.c_longlong]
    lib.ref_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.ref_full_set_seed(seed)
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return lib.ref_main(len(args), arr)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
