"""Sharding of one large compression job over the GPUs of a box.

The path shards by construction: the encoder cuts its input into independent chunks that each end
in a byte aligned sync marker (reference endstream(), src/deflator.c:609-654), so rank g can
compress the contiguous byte range [g*N/G, (g+1)*N/G) (rounded to the chunk size) on its own GPU
and the per-rank outputs concatenate into ONE valid DEFLATE stream.  Only the last rank ends its
part with DEFLT_END (BFINAL = 1); every other rank ends with DEFLT_FLUSH.

The only collective is one all_gather of 24 bytes per rank -- {compressed bytes, checksum of the
rank's uncompressed slice, slice length} -- after which every rank knows its output offset
(exclusive scan) and the whole-stream CRC-32 / Adler-32 (ordered combine).  It runs over NCCL on
the GPUs (NVLink / NVSwitch) or gloo on CPU for the tests; no payload byte crosses ranks unless
the caller asks for a gathered copy.

This module holds only host-side planning and the combine arithmetic; compression itself is
whatever codec object the caller drives (the product library on a GPU).
"""
from __future__ import annotations

from dataclasses import dataclass

ADLER_MOD = 65521
CRC_POLY = 0xEDB88320


@dataclass
class ShardPlan:
    rank: int
    world: int
    begin: int          # first uncompressed byte of this rank
    end: int            # one past the last
    last: bool          # this rank closes the stream (DEFLT_END), the others DEFLT_FLUSH


def plan(total_bytes: int, world: int, rank: int, chunk_bytes: int = 512 << 10) -> ShardPlan:
    """Contiguous, chunk aligned slice of rank `rank` (reference-independent: SURVEY 8e)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    nchunks = (total_bytes + chunk_bytes - 1) // chunk_bytes
    c0 = nchunks * rank // world
    c1 = nchunks * (rank + 1) // world
    begin = min(total_bytes, c0 * chunk_bytes)
    end = min(total_bytes, c1 * chunk_bytes)
    # the stream is closed by the last rank that owns any byte (or rank world-1 for empty input)
    last_owner = world - 1
    if total_bytes:
        for r in range(world - 1, -1, -1):
            if nchunks * r // world < nchunks * (r + 1) // world:
                last_owner = r
                break
    return ShardPlan(rank, world, begin, end, rank == last_owner)


def _gf2_mul(a: int, b: int) -> int:
    p = 0
    for i in range(32):
        if a & (0x80000000 >> i):
            p ^= b
        b = (b >> 1) ^ CRC_POLY if b & 1 else b >> 1
    return p


def crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    """crc(A || B) from the finalised crc(A), crc(B) and len(B) (64-bit lengths; the reference's
    crc32_ncombine, src/zstrm.c:1413-1443, takes a u32 length)."""
    xp = 0x80000000
    for _ in range(8):
        xp = (xp >> 1) ^ CRC_POLY if xp & 1 else xp >> 1
    m = 0x80000000
    while len2:
        if len2 & 1:
            m = _gf2_mul(m, xp)
        xp = _gf2_mul(xp, xp)
        len2 >>= 1
    return _gf2_mul(crc1, m) ^ crc2


def adler32_combine(ad1: int, ad2: int, len2: int) -> int:
    """adler(A || B) from adler(A), adler(B), len(B) (no counterpart in the reference)."""
    a1, b1 = ad1 & 0xFFFF, ad1 >> 16
    a2, b2 = ad2 & 0xFFFF, ad2 >> 16
    a = (a1 + a2 + ADLER_MOD - 1) % ADLER_MOD
    b = (b1 + b2 + (len2 % ADLER_MOD) * ((a1 + ADLER_MOD - 1) % ADLER_MOD)) % ADLER_MOD
    return (b << 16) | a


def combine(per_rank):
    """per_rank: list of (compressed_bytes, checksum, uncompressed_bytes, kind) in rank order,
    kind 'crc32' | 'adler32' | None.  Returns (offsets, total_compressed, checksum, total_bytes)."""
    offsets, total, n = [], 0, 0
    kind = per_rank[0][3] if per_rank else None
    acc = 0 if kind == "crc32" else 1
    for comp, ck, raw, _ in per_rank:
        offsets.append(total)
        total += comp
        if kind == "crc32":
            acc = crc32_combine(acc, ck, raw)
        elif kind == "adler32":
            acc = adler32_combine(acc, ck, raw)
        n += raw
    return offsets, total, acc, n


def exchange(dist, device, compressed_bytes: int, checksum: int, raw_bytes: int, kind="crc32"):
    """The collective: all_gather of {compressed bytes, checksum, length} (3 x int64 = 24 B per
    rank).  `dist` is torch.distributed (initialised: nccl on GPUs, gloo on CPU).  Returns what
    combine() returns plus this rank's offset."""
    import torch
    mine = torch.tensor([compressed_bytes, checksum, raw_bytes], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, mine)
    rows = [tuple(int(x) for x in v.tolist()) + (kind,) for v in allv]
    offsets, total, ck, n = combine(rows)
    return offsets[dist.get_rank()], offsets, total, ck, n
