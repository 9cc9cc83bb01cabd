"""apde_mvs_b200 -- B200-native (sm_100a) implementation of the APDe-MVS hot path.

The product is ``libapde.so`` (hand-written CUDA behind the C ABI of ``include/apde.h``) plus the C++ host side
(``csrc/host``: the ``APD`` problem class, the ``apd`` CLI).  This Python package is only the ctypes mirror of the C
ABI used by the tests, ``bench.py`` and ``run.py``; it contains no compute and no fallback: every call goes to the
CUDA library and raises if the library or a GPU is missing.
"""
from .binding import (  # noqa: F401
    ApdeError,
    Camera,
    Context,
    Params,
    Schedule,
    Timing,
    FIELD,
    STAGE,
    lib_path,
    load_library,
)
