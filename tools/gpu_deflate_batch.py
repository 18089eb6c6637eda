"""Quick GPU check of the batched deflate path: records -> one zlib stream each, device resident,
then back through the batched inflate.  usage: gpu_deflate_batch.py [n_distinct] [tile] [record bytes] [level]"""
import sys, pathlib, time, zlib
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import numpy as np, torch
from support import Corpus
from jdeflate_b200 import api

n_distinct = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rbytes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
level = int(sys.argv[4]) if len(sys.argv) > 4 else 6
jd = api.load(); c = Corpus()
recs = [c.json_record(i) for i in range(n_distinct)]
if rbytes:
    recs = [r[:rbytes] for r in recs]
count = n_distinct * tile
perm = np.random.RandomState(1).permutation(count) % n_distinct
lens = np.array([len(r) for r in recs], np.uint64)
src_off = np.zeros(n_distinct + 1, np.uint64); src_off[1:] = np.cumsum(lens)
L = lens[perm]
caps = L + L // np.uint64(64) + np.uint64(80)
items = np.zeros((count, 4), np.uint64)
items[:, 0] = src_off[perm]
items[:, 1] = np.concatenate([[0], np.cumsum(caps)[:-1]])
items[:, 2] = L
items[:, 3] = caps
total_in = int(L.sum()); total_cap = int(caps.sum())
src = torch.from_numpy(np.frombuffer(b"".join(recs), np.uint8).copy()).cuda()
out = torch.empty(total_cap, dtype=torch.uint8, device="cuda")
ditems = torch.from_numpy(items.view(np.int64)).cuda()
dres = torch.zeros((count, 4), dtype=torch.int64, device="cuda")
print("records", count, "mean bytes", total_in / count, "in GB", total_in / 1e9, flush=True)
jd.profile(True)
for it in range(4):
    if it == 1: jd.profile(True)
    torch.cuda.synchronize(); t = time.time()
    rc = jd.deflate_batch(src.data_ptr(), out.data_ptr(), ditems.data_ptr(), dres.data_ptr(), count, api.JDB200_ZLIB, level)
    torch.cuda.synchronize(); dt = time.time() - t
    print("rc", rc, "time ms %.2f" % (dt * 1e3), "GB/s in %.2f" % (total_in / dt / 1e9), "Mrec/s %.2f" % (count / dt / 1e6), flush=True)
prof = jd.profile_read(); jd.profile(False)
print("   per call ms:", {k: round(v[1] / 3, 3) for k, v in prof.items()}, flush=True)
res = dres.cpu().numpy().view(np.uint32).reshape(count, 8)
used = dres.cpu().numpy().view(np.uint64).reshape(count, 4)[:, 3]
print("status!=0:", int((res[:, 0] != 0).sum()), "compressed GB", used.sum() / 1e9, "ratio %.3f" % (total_in / used.sum()))
zsample = sum(len(zlib.compress(recs[i], level)) for i in range(min(n_distinct, 512))) / sum(len(recs[i]) for i in range(min(n_distinct, 512)))
print("zlib ratio (sample) %.3f" % (1 / zsample))
# back through the batched inflate
it2 = items.copy()
it2[:, 0] = items[:, 1]; it2[:, 2] = used
it2[:, 1] = np.concatenate([[0], np.cumsum(L)[:-1]]); it2[:, 3] = L
back = torch.empty(total_in, dtype=torch.uint8, device="cuda")
d2 = torch.from_numpy(it2.view(np.int64)).cuda()
r2 = torch.zeros((count, 4), dtype=torch.int64, device="cuda")
rc = jd.inflate_batch(out.data_ptr(), back.data_ptr(), d2.data_ptr(), r2.data_ptr(), count, api.JDB200_ZLIB)
torch.cuda.synchronize()
rr = r2.cpu().numpy().view(np.uint32).reshape(count, 8)
print("inflate rc", rc, "status!=0:", int((rr[:, 0] != 0).sum()), "zerror!=0:", int((rr[:, 2] != 0).sum()))
host = back.cpu().numpy()
bad = 0
for k in range(0, count, max(1, count // 300)):
    o, n = int(it2[k, 1]), int(it2[k, 3])
    if host[o:o + n].tobytes() != recs[perm[k]]: bad += 1
print("content mismatches (sampled):", bad)
