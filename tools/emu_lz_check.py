"""Emulator check of the LZ stage: tokens must reproduce the input exactly."""
import sys, pathlib, ctypes as C, time
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np
from support import Corpus, KIND_NAMES

lib = C.CDLL(str(R / "tests/simt/_build/libjdeflate_emu.so"))
lib.jdb_dev_alloc.restype = C.c_void_p; lib.jdb_dev_alloc.argtypes = [C.c_size_t]
lib.jdb_lz_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.jdb_lz_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p] + [C.c_uint32] * 6 + [C.c_void_p] * 4
SEG = 8192

def run(data: bytes, chunk=262144, rng=131072, good=16, nice=16, chain=48, lazy=1):
    n = len(data); nseg = (n + SEG - 1) // SEG
    npad = nseg * SEG + 64
    din = lib.jdb_dev_alloc(npad); C.memmove(din, data, n)
    prev = lib.jdb_dev_alloc(npad * 2); tok = lib.jdb_dev_alloc(npad * 4)
    ntok = lib.jdb_dev_alloc(nseg * 4 + 64); hist = lib.jdb_dev_alloc(nseg * 320 * 4)
    assert lib.jdb_lz_chain(din, n, chunk, rng, None, prev, None, None) == 0
    assert lib.jdb_lz_parse(din, n, chunk, None, prev, good, nice, chain, lazy, 0, 0, tok, ntok, hist, None) == 0
    prev_a = np.frombuffer((C.c_uint16 * n).from_address(prev), np.uint16)
    tok_a = np.frombuffer((C.c_uint32 * (nseg * SEG)).from_address(tok), np.uint32)
    ntok_a = np.frombuffer((C.c_uint32 * nseg).from_address(ntok), np.uint32)
    hist_a = np.frombuffer((C.c_uint32 * (nseg * 320)).from_address(hist), np.uint32).reshape(nseg, 320)
    return prev_a, tok_a, ntok_a, hist_a

def verify(data, chunk, tok_a, ntok_a, hist_a):
    out = bytearray(); nlit = nmatch = mbytes = 0
    for s in range(len(ntok_a)):
        h = np.zeros(320, np.int64)
        seg_start = len(out)
        for t in tok_a[s * SEG: s * SEG + ntok_a[s]]:
            t = int(t)
            if t & 0x80000000:
                ln = ((t >> 16) & 0xff) + 3; d = (t & 0x7fff) + 1
                chunk0 = seg_start // chunk * chunk
                assert len(out) - d >= chunk0, ("match crosses chunk start", s, len(out), d)
                for _ in range(ln): out.append(out[-d])
                nmatch += 1; mbytes += ln
            else:
                out.append(t); nlit += 1; h[t] += 1
        assert len(out) == min((s + 1) * SEG, len(data)), ("segment size", s, len(out))
        assert hist_a[s][:256].tolist() == h[:256].tolist(), "literal histogram"
        assert hist_a[s].sum() == ntok_a[s] + (hist_a[s][288:].sum()), "hist total"
    assert bytes(out) == data, "reconstruction"
    return nlit, nmatch, mbytes

if __name__ == "__main__":
    c = Corpus()
    for kind in range(5):
        for n in (1, 3, 100, 16384, 40000, 300000):
            d = c.fill(kind, n, offset=4242)
            t = time.time()
            prev_a, tok_a, ntok_a, hist_a = run(d)
            nlit, nmatch, mb = verify(d, 262144, tok_a, ntok_a, hist_a)
            print(KIND_NAMES[kind], n, "lits", nlit, "matches", nmatch, "avg len %.1f" % (mb / max(nmatch, 1)), "%.1fs" % (time.time() - t))
