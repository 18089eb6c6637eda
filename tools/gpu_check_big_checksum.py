"""GPU check: zstrm_crc32update / zstrm_adler32update on device buffers of more than 1, 2 and 4 GiB against zlib."""
import sys, pathlib, zlib
R = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "tests"))
import torch
from support import Corpus
from jdeflate_b200 import api
jd = api.load(); c = Corpus()
MIB = 1 << 20
tile = 64 * MIB
tb = c.fill(1, tile, offset=12345)
piece = torch.frombuffer(bytearray(tb), dtype=torch.uint8).cuda()
for n in (tile, (1 << 30) + 5 * MIB, (2 << 30) + 3, (4 << 30), (4 << 30) + 5 * MIB):
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    crc, ad = 0, 1
    for off in range(0, n, tile):
        k = min(tile, n - off)
        dev[off:off + k].copy_(piece[:k])
        crc = zlib.crc32(tb[:k], crc); ad = zlib.adler32(tb[:k], ad)
    torch.cuda.synchronize()
    g = jd.lib.zstrm_crc32update(0xFFFFFFFF, dev.data_ptr(), n) ^ 0xFFFFFFFF
    a = jd.lib.zstrm_adler32update(1, dev.data_ptr(), n)
    print(n, "crc", g == crc, "adler", a == ad, hex(g), hex(crc))
    del dev
