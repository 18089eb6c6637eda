"""Multi-GPU sharding logic on CPU: world_size-2 gloo processes each compress
their slice with the SIMT-emulator build (test infrastructure) of the very same library, exchange
24 bytes through the one collective of the path, and the concatenation must be a single valid
gzip-equivalent stream with the right CRC."""
import os
import socket
import zlib

import pytest

from jdeflate_b200 import shard


def test_plan_covers_everything():
    for total in (0, 1, 100, 262144, 262145, 10 * 262144 + 7, 1 << 30):
        for world in (1, 2, 3, 8):
            plans = [shard.plan(total, world, r) for r in range(world)]
            assert plans[0].begin == 0 and plans[-1].end == total
            for a, b in zip(plans, plans[1:]):
                assert a.end == b.begin
            assert sum(p.last for p in plans) == 1
            last = [p for p in plans if p.last][0]
            assert last.end == total and all(p.begin == p.end for p in plans[last.rank + 1:])
            for p in plans:
                assert p.begin % 262144 == 0 or p.begin == total


def test_combine_arithmetic(corpus):
    d = corpus.fill(5, 3 << 20)
    cuts = [0, 700001, 700001, 2 << 20, len(d)]
    rows_c, rows_a = [], []
    for a, b in zip(cuts, cuts[1:]):
        rows_c.append((10, zlib.crc32(d[a:b]), b - a, "crc32"))
        rows_a.append((10, zlib.adler32(d[a:b]), b - a, "adler32"))
    offs, total, crc, n = shard.combine(rows_c)
    assert (offs, total, crc, n) == ([0, 10, 20, 30], 40, zlib.crc32(d), len(d))
    assert shard.combine(rows_a)[2] == zlib.adler32(d)
    assert shard.crc32_combine(0x12345678, 0x9ABCDEF0, (1 << 35) + 3) == \
        shard.crc32_combine(shard.crc32_combine(0x12345678, 0, 1 << 35), 0x9ABCDEF0, 3)


def _worker(rank, world, port, total, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from pathlib import Path
        root = Path(__file__).resolve().parent.parent
        sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "tests"))
        from jdeflate_b200 import api
        from jdeflate_b200.build import emu_lib_path
        from support import Corpus
        lib = api.JDeflateLib(emu_lib_path())
        pl = shard.plan(total, world, rank, chunk_bytes=65536)
        data = Corpus().fill(5, total, offset=(4 << 20) - 100000)[pl.begin:pl.end]
        de = lib.deflator(6)
        comp = de.run(data, flush=api.DEFLT_END if pl.last else api.DEFLT_FLUSH) if (data or pl.last) else b""
        de.close()
        off, offsets, tot, crc, n = shard.exchange(dist, "cpu", len(comp), zlib.crc32(data), len(data))
        with open(os.path.join(tmp, f"part{rank}"), "wb") as f:
            f.write(comp)
        with open(os.path.join(tmp, f"meta{rank}"), "w") as f:
            f.write(f"{off} {tot} {crc} {n}")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip(tmp_path, corpus, emu):
    import torch.multiprocessing as mp
    total = 300000
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    parts = [(tmp_path / f"part{r}").read_bytes() for r in range(2)]
    metas = [[int(x) for x in (tmp_path / f"meta{r}").read_text().split()] for r in range(2)]
    data = corpus.fill(5, total, offset=(4 << 20) - 100000)
    assert metas[0][0] == 0 and metas[1][0] == len(parts[0])
    assert metas[0][1:] == metas[1][1:] == [len(parts[0]) + len(parts[1]), zlib.crc32(data), total]
    # the concatenation is ONE valid raw DEFLATE stream
    assert zlib.decompress(parts[0] + parts[1], -15) == data
