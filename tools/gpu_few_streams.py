"""A batch of FEW large third-party streams (zlib level 6, no sync markers) through jdb200_inflate_batch:
fewer streams than SMs run one thread block each (inflate_wide_kernel); JDB200_INFLATE_NARROW=1 gives the
one-warp-per-stream kernel for comparison.  usage: gpu_few_streams.py [streams] [MiB per stream] [kind]"""
import sys, pathlib, time, zlib, os, ctypes as C
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 5
jd = api.load(os.environ.get("JDB200_LIB")); c = Corpus(); n = mib << 20
datas = [c.fill(kind, n, offset=i * n) for i in range(ns)]
zs = [zlib.compress(d, 6) for d in datas]
items = (api.BatchItem * ns)(); so = 0
for i, z in enumerate(zs):
    items[i] = api.BatchItem(so, i * n, len(z), n); so += (len(z) + 15) & ~15
src = np.zeros(so, np.uint8)
for i, z in enumerate(zs): src[items[i].srcoffset: items[i].srcoffset + len(z)] = np.frombuffer(z, np.uint8)
dsrc = torch.from_numpy(src).cuda(); ddst = torch.empty(ns * n, dtype=torch.uint8, device="cuda")
ditems = torch.from_numpy(np.frombuffer(bytes(items), np.uint8).copy()).cuda()
dres = torch.zeros(ns * C.sizeof(api.BatchResult), dtype=torch.uint8, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    rc = jd.lib.jdb200_inflate_batch(dsrc.data_ptr(), ddst.data_ptr(), ditems.data_ptr(), dres.data_ptr(), ns, api.JDB200_ZLIB)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("rc", rc, "streams", ns, "x", mib, "MiB: %.1f ms, %.2f GB/s out" % (dt * 1e3, ns * n / dt / 1e9), flush=True)
res = np.frombuffer(dres.cpu().numpy().tobytes(), dtype=np.uint32).reshape(ns, -1)
back = ddst.cpu().numpy()
print("status != 0:", int((res[:, 0] != 0).sum()), "zerror != 0:", int((res[:, 2] != 0).sum()),
      "content equal:", all(back[i * n:(i + 1) * n].tobytes() == datas[i] for i in range(ns)))
