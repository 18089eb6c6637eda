"""Quick GPU check of the batched inflate path (BASELINE config 3 shape, scaled)."""
import sys, pathlib, time, zlib, ctypes as C
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api

n_distinct = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 16
import os
jd = api.load(os.environ.get("JDB200_LIB")); c = Corpus()
lib = jd.lib
lib.jdb200_inflate_batch.restype = C.c_int
lib.jdb200_inflate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
t = time.time()
recs = [c.json_record(i) for i in range(n_distinct)]
comp = [zlib.compress(r, 6) for r in recs]
print("records", n_distinct, "raw MB", sum(map(len, recs)) / 1e6, "comp MB", sum(map(len, comp)) / 1e6, "prep s", time.time() - t)
count = n_distinct * tile
perm = np.random.RandomState(1).permutation(count) % n_distinct
src_off = np.zeros(n_distinct + 1, np.uint64); src_off[1:] = np.cumsum([len(x) for x in comp])
items = np.zeros((count, 4), np.uint64)
dst = 0
for k, i in enumerate(perm):
    items[k] = (src_off[i], dst, len(comp[i]), len(recs[i])); dst += len(recs[i])
src = torch.from_numpy(np.frombuffer(b"".join(comp), np.uint8).copy()).cuda()
out = torch.empty(dst, dtype=torch.uint8, device="cuda")
ditems = torch.from_numpy(items.view(np.int64)).cuda()
dres = torch.zeros((count, 4), dtype=torch.int64, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    rc = lib.jdb200_inflate_batch(src.data_ptr(), out.data_ptr(), ditems.data_ptr(), dres.data_ptr(), count, 1)
    torch.cuda.synchronize(); dt = time.time() - t
    print("rc", rc, "batch", count, "out GB", dst / 1e9, "time ms", dt * 1e3, "GB/s out", dst / dt / 1e9)
res = dres.cpu().numpy().view(np.uint32).reshape(count, 8)
print("status!=0:", int((res[:, 0] != 0).sum()), "zerror!=0:", int((res[:, 2] != 0).sum()))
host = out.cpu().numpy()
bad = 0
for k in range(0, count, max(1, count // 200)):
    o, n = int(items[k, 1]), int(items[k, 3])
    if host[o:o + n].tobytes() != recs[perm[k]]: bad += 1
print("content mismatches (sampled):", bad)
