"""Time the decode of ONE third-party stream (zlib level 6, no sync markers: nothing to parallelise across chunks)
through the plain inflator with device buffers.  usage: gpu_single_stream.py [MiB] [corpus kind]"""
import sys, pathlib, time, zlib
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
import os
jd = api.load(os.environ.get("JDB200_LIB")); c = Corpus(); n = mib << 20
data = c.fill(kind, n, offset=0)
z = zlib.compressobj(6, zlib.DEFLATED, -15); raw = z.compress(data) + z.flush()
src = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
want = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for it in range(2):
    s = jd.inflator(); s.setsrc(src.data_ptr(), len(raw)); s.settgt(back.data_ptr(), n)
    if it == 1: jd.profile(True)
    torch.cuda.synchronize(); t = time.perf_counter()
    r = s.inflate(1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    if it == 1:
        print("kernels:", {k: (v if isinstance(v, (int, float)) else v) for k, v in jd.profile_read().items()}); jd.profile(False)
    s.close()
    print("kind", kind, "rc", r, "ms %.1f" % (dt * 1e3), "MB/s %.1f" % (n / dt / 1e6), "ratio %.2f" % (n / len(raw)), flush=True)
print("equal", bool(torch.equal(back, want)))
