import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import zlib, os, time, ctypes
import torch
from jdeflate_b200 import api
jd = api.load()
print("version", jd.version())
bad = 0
for n in [1,15,16,17,100,4096,5000,65536,70001,1<<20,(1<<24)+5]:
    for off in (0,1,7):
        data = os.urandom(n+off)[off:]
        c = jd.crc32(data); a = jd.adler32(data)
        if c != zlib.crc32(data) or a != zlib.adler32(data):
            bad += 1; print("MISMATCH", n, off)
print("host-pointer checksum mismatches:", bad)
x = torch.randint(0, 256, (1<<30,), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for which in ("crc32", "adler32"):
    fn = getattr(jd, which)
    fn(x.data_ptr(), x.numel())
    t = time.time(); 
    for _ in range(5): v = fn(x.data_ptr(), x.numel())
    dt = (time.time()-t)/5
    print(which, hex(v), "%.1f GB/s" % (x.numel()/dt/1e9))
h = x[:1<<26].cpu().numpy().tobytes()
print("dev crc ok", jd.crc32(x.data_ptr(), 1<<26) == zlib.crc32(h), jd.adler32(x.data_ptr(), 1<<26) == zlib.adler32(h))
