"""monocularsfm_b200 — B200-native (sm_100a) hot path of MonocularSfM.

The product is the C-ABI shared library ``libmsfm_b200.so`` (see ``include/msfm_b200.h``) built from
``monocularsfm_b200/csrc`` plus the C++ classes in ``monocularsfm_b200/host`` that keep the reference's
names.  This Python package is only the ctypes binding used by the tests, ``bench.py`` and the
``torch.distributed`` plumbing; it contains no compute and no CPU fallback.
"""
from ._ffi import (  # noqa: F401
    BAOptions,
    BAProblem,
    Context,
    MatchOptions,
    MsfmError,
    lib_path,
    load_library,
    exported_symbols,
)

__all__ = ["BAOptions", "BAProblem", "Context", "MatchOptions", "MsfmError", "lib_path", "load_library", "exported_symbols"]
