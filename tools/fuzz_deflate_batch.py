"""GPU / emulator fuzz of jdb200_deflate_batch: random record counts, sizes, kinds, levels, formats,
slot sizes and target capacities; every stream must decode bit-exactly through zlib, results must
carry the right sizes / Adler-32, records whose target is too small must report DEFLT_TGTEXHSTD."""
import os, sys, time, zlib, random, pathlib
R = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(R)); sys.path.insert(0, str(R / 'tests'))
from support import Corpus
from jdeflate_b200 import api
lib = api.JDeflateLib(sys.argv[1]); c = Corpus()
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
ncase = int(sys.argv[3]) if len(sys.argv) > 3 else 60
bad = 0; t = time.time(); nrec = 0
for i in range(ncase):
    lvl = rnd.choice([0, 1, 3, 6, 6, 9]); fmt = rnd.choice([api.JDB200_RAW, api.JDB200_ZLIB])
    cnt = rnd.choice([1, 2, rnd.randrange(3, 40), rnd.randrange(40, 400)])
    big = rnd.choice([300, 3000, 20000, 200000])
    recs = []
    for _ in range(cnt):
        n = rnd.choice([0, 1, rnd.randrange(0, 64), rnd.randrange(0, big)])
        recs.append(c.fill(rnd.randrange(6), n, offset=rnd.randrange(0, 1 << 24)) if n else b"")
    for k in ("JDB200_RECORD_CHUNK_KIB", "JDB200_BATCH_MIB"):
        os.environ.pop(k, None)
    if rnd.random() < 0.4: os.environ["JDB200_RECORD_CHUNK_KIB"] = str(rnd.choice([16, 32, 64, 512]))
    if rnd.random() < 0.4: os.environ["JDB200_BATCH_MIB"] = str(rnd.choice([1, 2, 8]))
    caps = [len(r) + len(r) // 64 + 80 for r in recs]
    small = set()
    if rnd.random() < 0.3:
        for k in rnd.sample(range(cnt), max(1, cnt // 10)):
            caps[k] = rnd.randrange(0, 5); small.add(k)        # an empty record needs 5 bytes
    outs, res = lib.deflate_batch_bytes(recs, caps=caps, fmt=fmt, level=lvl)
    ok = True
    for k, (r, z, q) in enumerate(zip(recs, outs, res)):
        if k in small:
            ok = ok and (q.status, q.tgtused, q.srcused) == (2, 0, 0)
            continue
        try:
            back = zlib.decompress(z) if fmt == api.JDB200_ZLIB else zlib.decompress(z, -15)
        except zlib.error:
            back = None
        ok = ok and back == r and (q.status, q.srcused, q.tgtused) == (0, len(r), len(z))
        if fmt == api.JDB200_ZLIB: ok = ok and q.checksum == zlib.adler32(r)
    nrec += cnt
    if not ok:
        bad += 1; print("BAD case", i, "L%d" % lvl, "fmt", fmt, "records", cnt, dict((k, os.environ.get(k)) for k in ("JDB200_RECORD_CHUNK_KIB", "JDB200_BATCH_MIB")))
print("deflate batch fuzz cases", ncase, "records", nrec, "bad", bad, "%.1fs" % (time.time() - t))
