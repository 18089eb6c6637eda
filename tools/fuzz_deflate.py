"""GPU / emulator fuzz of the encoder: random sizes, kinds, levels, window splits; every stream must
decode bit-exactly through zlib and the oracle, size within 3 % of the reference encoder (oracle)."""
import sys, time, zlib, random, pathlib
R = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(R)); sys.path.insert(0, str(R / 'tests'))
from support import Corpus, Oracle, KIND_NAMES
from jdeflate_b200 import api
lib = api.JDeflateLib(sys.argv[1]); o = Oracle(); c = Corpus()
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 11)
ncase = int(sys.argv[3]) if len(sys.argv) > 3 else 200
bad = 0; t = time.time(); worst = 0.0
for i in range(ncase):
    kind = rnd.randrange(6); lvl = rnd.choice([0, 1, 2, 3, 4, 5, 6, 6, 6, 7, 8, 9])
    n = rnd.choice([rnd.randrange(0, 300), rnd.randrange(300, 70000), rnd.randrange(70000, 3000000)])
    d = c.fill(kind, n, offset=rnd.randrange(0, 1 << 24)) if n else b""
    kw = {}
    if rnd.random() < 0.3: kw["window"] = rnd.choice([7, 100, 4096, 100000])
    if rnd.random() < 0.3 and n: kw["feed"] = rnd.choice([1000, 65536, 300000])
    z = lib.deflate_bytes(d, lvl, **kw)
    ok = zlib.decompress(z, -15) == d
    st, err, out, used = o.inflate(z, n + 1)
    ok = ok and (st, err, used) == (0, 0, len(z)) and out == d
    ratio = 0.0
    if n >= 70000:
        ref = len(o.deflate(d, lvl)); ratio = len(z) / ref - 1; worst = max(worst, ratio)
        ok = ok and len(z) <= 1.03 * ref + 16
    if not ok:
        bad += 1; print("BAD", KIND_NAMES[kind], n, "L%d" % lvl, kw, "ratio %+.2f%%" % (100 * ratio))
print("deflate fuzz cases", ncase, "bad", bad, "worst size vs reference %+.2f%%" % (100 * worst), "%.1fs" % (time.time() - t))
