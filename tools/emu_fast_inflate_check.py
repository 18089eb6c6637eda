"""Emulator / GPU check: the fast batch inflate path gives the same per-record results as the general decoder."""
import sys, pathlib, os, zlib, random, subprocess, json
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
from support import Corpus, zlib_raw, Oracle
from jdeflate_b200 import api
from jdeflate_b200.build import emu_lib_path

def run(libpath, fmt, streams, caps):
    jd = api.JDeflateLib(libpath)
    outs, res = jd.inflate_batch_bytes(streams, caps, fmt=fmt)
    return [(o, r.status, r.error, r.zerror, r.checksum, r.srcused, r.tgtused) for o, r in zip(outs, res)]

if __name__ == "__main__":
    libpath = sys.argv[1] if len(sys.argv) > 1 else str(emu_lib_path())
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    c = Corpus(); o = Oracle(); rnd = random.Random(5)
    recs = [c.json_record(i) for i in range(n)]
    recs += [c.fill(k, rnd.randint(1, 40000), offset=rnd.randint(0, 1 << 22)) for k in (0, 1, 2, 3, 4) for _ in range(4)]
    recs += [b"", b"a", b"ab" * 5000]
    for fmt in (api.JDB200_ZLIB, api.JDB200_RAW):
        streams = []
        for i, r in enumerate(recs):
            lvl = (1, 6, 9)[i % 3]
            if fmt == api.JDB200_ZLIB:
                streams.append(zlib.compress(r, lvl))
            else:
                streams.append(zlib_raw(r, lvl) if i % 2 else o.deflate(r, (0, 1, 6)[i % 3]))
        caps = [len(r) for r in recs]
        # damage a few
        for i in (3, 17, 31):
            b = bytearray(streams[i]); b[len(b) // 2] ^= 0x21; streams[i] = bytes(b)
        streams[5] = streams[5][: len(streams[5]) // 2]
        caps[7] = max(1, caps[7] // 2)
        streams[9] = streams[9] + b"trailing junk"
        os.environ.pop("JDB200_NO_FAST_INFLATE", None)
        os.environ["JDB200_FAST_INFLATE"] = "1"
        fast = run(libpath, fmt, streams, caps)
        os.environ["JDB200_NO_FAST_INFLATE"] = "1"
        slow = run(libpath, fmt, streams, caps)
        bad = 0
        for i, (f, s) in enumerate(zip(fast, slow)):
            same = f[1:] == s[1:] and (f[0] == s[0] or s[1] != 0)
            if not same:
                bad += 1
                print("MISMATCH fmt", fmt, "rec", i, "fast", f[1:], "slow", s[1:])
            if s[1] == 0 and s[3] == 0 and i not in (3, 17, 31):
                assert s[0] == recs[i][: caps[i]], i
        print("fmt", fmt, "records", len(recs), "mismatches", bad, "ok-status", sum(1 for s in slow if s[1] == 0))
