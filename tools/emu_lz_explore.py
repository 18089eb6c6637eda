"""Emulator-only exploration of the LZ stage: compressed size against the reference and
chain steps per input byte (the work the match search does), per corpus kind.

    python tools/emu_lz_explore.py [level] [KiB] [kinds]      e.g.  6 1024 0,1,2,4,5

Environment switches of lz.cu (JDB_LZ_*) are read when the emulator library launches, so
variants are compared by setting them around this script.
"""
import sys, pathlib, ctypes as C, time, zlib, os
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np
from support import Corpus, KIND_NAMES, Oracle
from jdeflate_b200.build import build_emu

lib = C.CDLL(str(build_emu()))


class Cfg(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("level", "fixedonly", "good", "nice", "chain", "lazy", "chunk_bytes",
                                          "block_segs", "chain_range", "final", "dict_region", "dict_pad")] + \
               [("chunk_len", C.c_void_p), ("wrap_head", C.c_uint32), ("wrap_tail", C.c_uint32)]


lib.jdb_dev_alloc.restype = C.c_void_p; lib.jdb_dev_alloc.argtypes = [C.c_size_t]
lib.jdb_dev_free.argtypes = [C.c_void_p]
lib.jdb_deflate_workspace_bytes.restype = C.c_size_t; lib.jdb_deflate_workspace_bytes.argtypes = [C.c_uint64, C.POINTER(Cfg)]
lib.jdb_deflate_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Cfg), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
PARAMS = {1: (8, 4, 2, 0), 2: (8, 8, 8, 0), 3: (8, 16, 16, 0), 4: (8, 32, 32, 0), 5: (8, 64, 128, 0), 6: (16, 16, 48, 1),
          7: (32, 64, 128, 1), 8: (64, 128, 320, 1), 9: (192, 256, 512, 1)}


def deflate(data, level=6, chunk=524288, block_segs=8, steps=None):
    g, n_, c, l = PARAMS[level]
    cfg = Cfg(level, 0, g, n_, c, l, chunk, block_segs, 0, 1, 0, 0, None, 0, 0)
    n = len(data)
    wb = lib.jdb_deflate_workspace_bytes(n, C.byref(cfg)); assert wb
    work = lib.jdb_dev_alloc(wb); din = lib.jdb_dev_alloc(n + 64); C.memmove(din, data, n)
    if steps is not None:
        C.c_void_p.in_dll(lib, "jdb_emu_lz_steps").value = steps.ctypes.data
    out = C.c_void_p(); tot = C.c_void_p()
    rc = lib.jdb_deflate_run(din, n, C.byref(cfg), work, C.byref(out), C.byref(tot), None); assert rc == 0, rc
    total = C.c_uint64.from_address(tot.value).value
    res = C.string_at(out.value, total)
    lib.jdb_dev_free(work); lib.jdb_dev_free(din)
    C.c_void_p.in_dll(lib, "jdb_emu_lz_steps").value = None
    return res


if __name__ == "__main__":
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    kib = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    kinds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 4, 5]
    c = Corpus(); o = Oracle()
    n = kib << 10
    tot_ours = tot_ref = 0
    for kind in kinds:
        d = c.fill(kind, n, offset=7 << 20)
        steps = np.zeros(n, np.uint16)
        t = time.time()
        z = deflate(d, level, steps=steps)
        dt = time.time() - t
        assert zlib.decompress(z, -15) == d
        ref = len(o.deflate(d, level))
        tot_ours += len(z); tot_ref += ref
        print("%-7s L%d ours %8d ref %8d delta %+6.2f%%  steps/byte %6.2f  searched %5.1f%%  (%.0fs)" % (
            KIND_NAMES[kind], level, len(z), ref, 100.0 * (len(z) - ref) / ref, steps.mean(),
            100.0 * (steps > 0).mean(), dt), flush=True)
    print("total   ours %d ref %d delta %+.2f%%" % (tot_ours, tot_ref, 100.0 * (tot_ours - tot_ref) / tot_ref))
