"""jdeflate_b200 -- B200-native DEFLATE / zlib / gzip codec behind the jdeflate C API.

The product is ``jdeflate_b200/lib/libjdeflate.so`` (C99 host layer + hand-written
sm_100a CUDA kernels, see ``csrc/``).  This package is only the thin ctypes mirror
of that C ABI used by the tests and ``bench.py``; it contains no codec logic and
no CPU fallback: importing :mod:`jdeflate_b200.api` fails loudly when the shared
library has not been built, and every constructor fails when no CUDA device is
present.
"""
from .build import build_cuda, build_oracle, lib_path  # noqa: F401

__all__ = ["build_cuda", "build_oracle", "lib_path"]
__version__ = "0.4.0+b200.r1"
