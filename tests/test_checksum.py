"""crc32 / adler32 parity (reference zstrm_crc32update / zstrm_adler32update / crc32_ncombine,
src/zstrm.c:1346-1526).  Oracle: zlib == oracle/jd_oracle.c == golden vectors; bit exact."""
import zlib

import pytest


SIZES = [1, 2, 3, 15, 16, 17, 255, 4095, 4096, 4097, 5551, 5552, 5553, 65520, 65521, 65522, 131071, 300001]


@pytest.mark.parametrize("kind", [0, 2, 3])
def test_checksums_match_oracle(lib, oracle, corpus, kind):
    for n in SIZES:
        d = corpus.fill(kind, n, offset=n)
        assert lib.crc32(d) == oracle.crc32(d) == zlib.crc32(d), n
        assert lib.adler32(d) == oracle.adler32(d) == zlib.adler32(d), n


def test_checksums_golden(lib, corpus, golden):
    for e in golden["checksums"]:
        if e["n"] > 100000 and e["kind"] not in (0, 3):
            continue
        d = corpus.fill(e["kind"], e["n"], offset=e["offset"])
        assert lib.crc32(d) == e["crc"]
        assert lib.adler32(d) == e["adler"]


def test_worst_case_bytes(lib):
    # all 0xff maximises the Adler sums (the size class the reference's C fallback gets wrong)
    for n in (5552, 65521, 70000, 262144 + 4095):
        d = b"\xff" * n
        assert lib.adler32(d) == zlib.adler32(d)
        assert lib.crc32(d) == zlib.crc32(d)


def test_incremental_register_semantics(lib, corpus):
    d = corpus.fill(1, 100003, offset=3)
    for cut in (1, 4096, 50000, 100002):
        assert lib.crc32(d[cut:], value=lib.crc32(d[:cut])) == zlib.crc32(d)
        assert lib.adler32(d[cut:], value=lib.adler32(d[:cut])) == zlib.adler32(d)
    # raw entry point: un-finalised register in, un-finalised register out
    reg = lib.lib.zstrm_crc32update(0xFFFFFFFF, d, len(d))
    assert reg ^ 0xFFFFFFFF == zlib.crc32(d)


def test_crc_combine(lib, oracle, corpus, golden):
    for e in golden["crc_combine"]:
        assert lib.crc32_combine(e["c1"], e["c2"], e["len2"]) == e["crc"]
        assert lib.lib.crc32_ncombine(e["c1"], e["c2"], e["len2"]) == e["crc"]
    a, b = corpus.fill(0, 12345), corpus.fill(3, 54321)
    assert lib.crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    # a checksum of checksums: combining per-piece CRCs in order equals the CRC of the whole
    d = corpus.fill(5, 1 << 20)
    acc = 0
    for i in range(0, len(d), 100000):
        piece = d[i:i + 100000]
        acc = lib.crc32_combine(acc, lib.crc32(piece), len(piece))
    assert acc == zlib.crc32(d)
    # 64-bit lengths (the reference takes a u32 length, src/zstrm.c:1428)
    assert lib.crc32_combine(0x12345678, 0x9abcdef0, (1 << 33) + 5) == oracle.crc32_combine(0x12345678, 0x9abcdef0, (1 << 33) + 5)


def test_wide_kernel_sizes(lib, corpus, monkeypatch):
    """Inputs of 16 MiB and more take ck_wide_kernel (1024 threads, bank-replicated row table);
    the threshold is lowered here so every span / padding shape of it is exercised at sizes zlib
    finishes quickly: fewer vectors than threads, ragged first rows, the 4-row unrolled loop and
    its remainder, unaligned heads and tails."""
    monkeypatch.setenv("JDB200_CK_WIDE_MIN_KIB", "1")
    big = corpus.fill(5, 3 << 20, offset=(4 << 20) - (1 << 20))
    for off, n in ((0, 1024), (1, 1040), (3, 16384 + 17), (0, 16384 * 5), (5, 16384 * 9 + 4001),
                   (0, 131072), (7, 262144 + 33), (2, 1 << 20), (9, (3 << 20) - 9)):
        d = big[off:off + n]
        assert lib.crc32(d) == zlib.crc32(d), (off, n)
        assert lib.adler32(d) == zlib.adler32(d), (off, n)
    ff = b"\xff" * ((2 << 20) + 5)
    assert lib.adler32(ff) == zlib.adler32(ff) and lib.crc32(ff) == zlib.crc32(ff)
    d = big[3:3 + (1 << 20) + 77]
    assert lib.crc32(d, value=12345) == zlib.crc32(d, 12345)
    assert lib.adler32(d) == zlib.adler32(d)
    # both registers in one pass (raw container with ZSTRM_DOCRC | ZSTRM_DOADLER)
    from jdeflate_b200 import api
    from test_zstrm import compress
    d = big[11:11 + (1 << 20) + 300]
    _, (crc, adler, total) = compress(lib, d, api.ZSTRM_DFLT, piece=len(d), flags=api.ZSTRM_DOCRC | api.ZSTRM_DOADLER)
    assert (crc, adler, total) == (zlib.crc32(d), zlib.adler32(d), len(d))
