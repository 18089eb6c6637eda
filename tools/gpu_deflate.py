"""Quick GPU check of the deflate pipeline through the C ABI (device-resident buffers)."""
import sys, pathlib, time, zlib, os, ctypes as C
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus, KIND_NAMES, Oracle
from jdeflate_b200 import api

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6]
kinds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2, 3, 5]
jd = api.load(os.environ.get('JDB200_LIB')); c = Corpus(); o = Oracle()
n = mib << 20
for kind in kinds:
    host = np.empty(n, np.uint8)
    c.fill_into(kind, host.ctypes.data, n)
    src = torch.from_numpy(host).cuda()
    out = torch.empty(n + n // 8 + 65536, dtype=torch.uint8, device="cuda")
    for lvl in levels:
        best = 1e9
        jd.profile(True)
        for it in range(4):
            d = jd.deflator(lvl)
            d.setsrc(src.data_ptr(), n)
            d.settgt(out.data_ptr(), out.numel())
            torch.cuda.synchronize(); t = time.time()
            r = d.deflate(api.DEFLT_END)
            torch.cuda.synchronize(); dt = time.time() - t
            best = min(best, dt)
            produced = d.tgtend(); consumed = d.srcend()
            d.close()
        prof = jd.profile_read(); jd.profile(False)
        print("   per call ms:", {k: round(v[1] / 4, 3) for k, v in prof.items()}, flush=True)
        comp = out[:produced].cpu().numpy().tobytes()
        ok = zlib.decompress(comp, -15) == host.tobytes()
        sample = host[: 8 << 20].tobytes()
        ref = len(o.deflate(sample, lvl)) * (n / len(sample))
        print(KIND_NAMES[kind], "L%d" % lvl, "rc", r, "consumed", consumed == n, "produced", produced,
              "ratio %.3f" % (n / produced), "vs ref(8MiB sample) %+.2f%%" % (100 * (produced - ref) / ref),
              "roundtrip", ok, "best ms %.2f" % (best * 1e3), "GB/s %.2f" % (n / best / 1e9), flush=True)
