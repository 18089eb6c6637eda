"""Emulator check of the whole deflate pipeline (jdb_deflate_run) against zlib and the reference."""
import sys, pathlib, ctypes as C, time, zlib
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
from support import Corpus, KIND_NAMES, Oracle

lib = C.CDLL(str(R / "tests/simt/_build/libjdeflate_emu.so"))
class Cfg(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("level","fixedonly","good","nice","chain","lazy","chunk_bytes","block_segs","chain_range","final","dict_region","dict_pad")]
lib.jdb_dev_alloc.restype = C.c_void_p; lib.jdb_dev_alloc.argtypes = [C.c_size_t]
lib.jdb_dev_free.argtypes = [C.c_void_p]
lib.jdb_deflate_workspace_bytes.restype = C.c_size_t; lib.jdb_deflate_workspace_bytes.argtypes = [C.c_uint64, C.POINTER(Cfg)]
lib.jdb_deflate_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Cfg), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
PARAMS = {1:(8,4,2,0),2:(8,8,8,0),3:(8,16,16,0),4:(8,32,32,0),5:(8,64,128,0),6:(16,16,48,1),7:(32,64,128,1),8:(64,128,320,1),9:(192,256,512,1)}

def deflate(data, level=6, chunk=262144, block_segs=4, final=1, fixedonly=0, chain_range=0):
    g,n_,c,l = PARAMS.get(level,(0,0,0,0))
    cfg = Cfg(level, fixedonly, g, n_, c, l, chunk, block_segs, chain_range, final, 0, 0)
    n = len(data)
    wb = lib.jdb_deflate_workspace_bytes(n, C.byref(cfg)); assert wb
    work = lib.jdb_dev_alloc(wb); din = lib.jdb_dev_alloc(n + 64); C.memmove(din, data, n)
    out = C.c_void_p(); tot = C.c_void_p()
    rc = lib.jdb_deflate_run(din, n, C.byref(cfg), work, C.byref(out), C.byref(tot), None); assert rc == 0, rc
    total = C.c_uint64.from_address(tot.value).value
    res = C.string_at(out.value, total)
    lib.jdb_dev_free(work); lib.jdb_dev_free(din)
    return res

if __name__ == "__main__":
    c = Corpus(); o = Oracle()
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 5, 300, 16384, 70000, 600000]
    for kind in range(5):
        for n in sizes:
            d = c.fill(kind, n, offset=999)
            for lvl in (0, 1, 6, 9):
                t = time.time()
                z = deflate(d, lvl)
                try:
                    back = zlib.decompress(z, -15)
                except Exception as e:
                    print("FAIL zlib", KIND_NAMES[kind], n, lvl, e); continue
                st, err, back2, used = o.inflate(z, n + 10)
                ok = back == d and back2 == d and st == 0 and used == len(z)
                refsz = len(o.deflate(d, lvl))
                print(("ok  " if ok else "BAD "), KIND_NAMES[kind], n, "L%d" % lvl, "ours", len(z), "ref", refsz, "delta %+.2f%%" % (100.0 * (len(z) - refsz) / max(refsz, 1)), "%.1fs" % (time.time() - t))
