"""gunzip of ONE ordinary .gz (written by zlib, no sync markers) through zstrm_inflate with host buffers and an
8 MiB source callback -- the drop-in case of a caller that only swaps the library.  usage: gpu_gunzip.py [MiB] [kind]"""
import sys, pathlib, time, zlib, os, ctypes as C
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np
from support import Corpus
from jdeflate_b200 import api
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 5
jd = api.load(os.environ.get("JDB200_LIB")); c = Corpus(); n = mib << 20
data = c.fill(kind, n, offset=0)
co = zlib.compressobj(6, zlib.DEFLATED, 31); gz = co.compress(data) + co.flush()
gzb = np.frombuffer(gz, np.uint8).copy(); piece = 8 << 20
back = np.empty(piece, np.uint8)
rd = {"p": 0}
def source(buf, size, user):
    k = min(size, piece, len(gz) - rd["p"])
    C.memmove(buf, gzb.ctypes.data + rd["p"], k); rd["p"] += k
    return k
icb = api.IFN(source)
zi = jd.lib.zstrm_create(api.ZSTRM_INFLATE | api.ZSTRM_GZIP, 0, None)
for it in range(3):
    jd.lib.zstrm_reset(zi); rd["p"] = 0
    jd.lib.zstrm_setsourcefn(zi, icb, None)
    total, crc = 0, 0
    t = time.perf_counter()
    while True:
        got = jd.lib.zstrm_inflate(zi, back.ctypes.data, piece)
        if got <= 0: break
        if it == 0: crc = zlib.crc32(back[:got], crc)
        total += got
    dt = time.perf_counter() - t
    print("gunzip %d MiB kind %d: error %d total ok %s %s %.1f ms %.3f GB/s" % (mib, kind, zi.contents.error, total == n,
          ("crc ok %s" % (crc == zlib.crc32(data))) if it == 0 else "", dt * 1e3, n / dt / 1e9), flush=True)
jd.lib.zstrm_destroy(zi)
t = time.perf_counter(); zlib.decompress(gz, 31); print("zlib on one host core: %.3f GB/s" % (n / (time.perf_counter() - t) / 1e9))
