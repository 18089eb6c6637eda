"""GPU: compress N MiB of the mixed corpus, then decode OUR OWN single stream through the plain
inflator (chunk-parallel path) and time it."""
import sys, pathlib, time, zlib
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
jd = api.load(); c = Corpus(); n = mib << 20
host = np.empty(n, np.uint8)
for off in range(0, n, 64 << 20):
    c.fill_into(5, host.ctypes.data + off, min(64 << 20, n - off), offset=off)
src = torch.from_numpy(host).cuda()
out = torch.empty(n + n // 8 + 65536, dtype=torch.uint8, device="cuda")
d = jd.deflator(6); d.setsrc(src.data_ptr(), n); d.settgt(out.data_ptr(), out.numel())
assert d.deflate(api.DEFLT_END) == api.OK
produced = d.tgtend(); d.close()
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for it in range(3):
    back.zero_()
    jd.profile(True)
    s = jd.inflator()
    s.setsrc(out.data_ptr(), produced); s.settgt(back.data_ptr(), n)
    torch.cuda.synchronize(); t = time.time()
    r = s.inflate(1)
    torch.cuda.synchronize(); dt = time.time() - t
    ok = r == api.OK and s.tgtend() == n and s.srcend() == produced
    s.close()
    prof = jd.profile_read(); jd.profile(False)
    print("inflate own stream: rc", r, "ok", ok, "equal", bool(torch.equal(back, src)), "ms %.1f" % (dt * 1e3), "GB/s %.2f" % (n / dt / 1e9),
          {k: (v[0], round(v[1], 2)) for k, v in prof.items() if "inflate" in k or "marker" in k}, flush=True)
