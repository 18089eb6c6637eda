"""Time the chunk-parallel decode of one large stream of ours through the plain inflator (device buffers)
and show where the time goes.  usage: gpu_own_stream.py [MiB]"""
import sys, pathlib, time
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
jd = api.load(); c = Corpus(); n = mib << 20
host = np.empty(n, np.uint8); c.fill_into(5, host.ctypes.data, n)
src = torch.from_numpy(host).cuda()
comp = torch.empty(n + n // 8 + 65536, dtype=torch.uint8, device="cuda")
d = jd.deflator(6); d.setsrc(src.data_ptr(), n); d.settgt(comp.data_ptr(), comp.numel())
assert d.deflate(api.DEFLT_END) == api.OK; clen = d.tgtend(); d.close()
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for it in range(3):
    jd.profile(True)
    s = jd.inflator(); s.setsrc(comp.data_ptr(), clen); s.settgt(back.data_ptr(), n)
    torch.cuda.synchronize(); t = time.perf_counter()
    r = s.inflate(1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    s.close()
    prof = jd.profile_read(); jd.profile(False)
    print("rc", r, "ms %.1f" % (dt * 1e3), "GB/s %.2f" % (n / dt / 1e9), {k: (v[0], round(v[1], 2)) for k, v in prof.items()}, flush=True)
print("equal", bool(torch.equal(back, src)))
