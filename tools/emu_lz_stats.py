"""Emulator-only: chain steps per position of lz_kernel (the work the match search does)."""
import sys, pathlib, ctypes as C
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np
from support import Corpus, KIND_NAMES
from jdeflate_b200.build import build_emu
lib = C.CDLL(str(build_emu()))
lib.jdb_dev_alloc.restype = C.c_void_p; lib.jdb_dev_alloc.argtypes = [C.c_size_t]
lib.jdb_lz_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
lib.jdb_lz_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p] + [C.c_uint32] * 6 + [C.c_void_p] * 4
SEG = 16384
c = Corpus()
n = 1 << 20
for kind in (0, 1, 2, 4):
    for (good, nice, chain, lazy) in ((16, 16, 48, 1), (192, 256, 512, 1)):
        data = c.fill(kind, n, offset=1 << 22)
        nseg = n // SEG
        din = lib.jdb_dev_alloc(n + 64); C.memmove(din, data, n)
        prev = lib.jdb_dev_alloc(n * 2 + 128); tok = lib.jdb_dev_alloc(n * 4 + 256)
        ntok = lib.jdb_dev_alloc(nseg * 4 + 64); hist = lib.jdb_dev_alloc(nseg * 320 * 4)
        iters = np.zeros(nseg * 1024, np.uint32); steps = np.zeros(n, np.uint8)
        C.c_void_p.in_dll(lib, "jdb_emu_lz_iters").value = iters.ctypes.data
        C.c_void_p.in_dll(lib, "jdb_emu_lz_steps").value = steps.ctypes.data
        assert lib.jdb_lz_chain(din, n, 262144, 262144, prev, None) == 0
        assert lib.jdb_lz_parse(din, n, 262144, prev, good, nice, chain, lazy, 0, 0, tok, ntok, hist, None) == 0
        st = steps.reshape(nseg, 16, 32, 32)          # seg, k, warp, lane  (p = tid + k*1024)
        sum_of_max = st.max(axis=3).sum(axis=1)       # nested loops: per warp sum over k of max over lanes
        print(KIND_NAMES[kind], "chain", chain, "chain steps per position: mean %.2f" % steps.mean(), "p50", np.percentile(steps, 50), "p90", np.percentile(steps, 90),
              "| per warp, if searched one position per lane at a time: sum over 16 rounds of the max over lanes = %.0f steps" % sum_of_max.mean())
