"""Raw DEFLATE encode parity (reference deflator_deflate, src/deflator.c:690-786).

Byte-identical output is not the goal of a chunk-parallel encoder; parity is:
every stream decodes bit-exactly through the reference's decoder (oracle restatement, and the
compiled reference when present) AND through zlib; the compressed size stays within 3 % of the
reference encoder at the same level (oracle.deflate is byte-identical to the reference, see
test_oracle.py); status codes and the streaming contract match."""
import zlib

import pytest

from jdeflate_b200 import api
from support import KIND_NAMES

RATIO_TOLERANCE = 1.03      # BASELINE.json north_star: within 3 percent of jdeflate at the same level


def check_stream(oracle, z, d):
    assert zlib.decompress(z, -15) == d
    st, err, out, used = oracle.inflate(z, len(d) + 1)
    assert (st, err, used) == (api.OK, 0, len(z)) and out == d


@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_roundtrip_and_ratio(lib, oracle, corpus, kind, level):
    for n in (1, 2, 3, 4, 5, 258, 4095, 16384, 16385, 70000, 300000):
        d = corpus.fill(kind, n, offset=3 * n)
        z = lib.deflate_bytes(d, level)
        check_stream(oracle, z, d)
        if n >= 70000:
            ref = len(oracle.deflate(d, level))
            assert len(z) <= RATIO_TOLERANCE * ref + 16, (KIND_NAMES[kind], n, level, len(z), ref)


@pytest.mark.parametrize("level", [2, 3, 4, 5, 7, 8])
def test_other_levels(lib, oracle, corpus, level):
    d = corpus.fill(4, 120000, offset=11)
    z = lib.deflate_bytes(d, level)
    check_stream(oracle, z, d)
    assert len(z) <= RATIO_TOLERANCE * len(oracle.deflate(d, level)) + 16


def test_ratio_at_one_mib_mixed(lib, oracle, corpus):
    """1 MiB straddling a TEXT -> BINARY segment boundary of the mixed corpus (BASELINE config 2 shape)."""
    d = corpus.fill(5, 1 << 20, offset=(4 << 20) - (1 << 19))
    z = lib.deflate_bytes(d, 6)
    check_stream(oracle, z, d)
    assert len(z) <= RATIO_TOLERANCE * len(oracle.deflate(d, 6))


def test_empty_input(lib, oracle):
    z = lib.deflate_bytes(b"", 6)
    assert z == bytes([1, 0, 0, 0xff, 0xff])           # endstream(): src/deflator.c:609-654
    check_stream(oracle, z, b"")


def test_fixed_codes_flag(lib, oracle, corpus):
    d = corpus.fill(0, 50000)
    z = lib.deflate_bytes(d, 6, flags=api.DEFLT_FIXEDCODES)
    check_stream(oracle, z, d)
    # only fixed (BTYPE 01) or stored blocks: the first block header says so
    assert (z[0] >> 1) & 3 == 1


def test_incompressible_is_stored(lib, oracle, corpus):
    d = corpus.fill(3, 100000)
    z = lib.deflate_bytes(d, 6)
    check_stream(oracle, z, d)
    assert len(z) <= len(d) + 5 * (len(d) // 65535 + 2) + 5     # DESIGN.md deviation 6
    assert len(z) < len(oracle.deflate(d, 6))


def test_streaming_windows(lib, oracle, corpus):
    """Any split of source pieces and target windows produces a valid stream
    (reference resumability, src/deflator.c:96-109)."""
    d = corpus.fill(1, 90000, offset=1)
    whole = lib.deflate_bytes(d, 6)
    for feed, window in ((1000, None), (None, 7), (4096, 100), (1, 65536)):
        if feed == 1:
            dd = d[:3000]
            z = lib.deflate_bytes(dd, 6, feed=feed, window=window)
            check_stream(oracle, z, dd)
            continue
        z = lib.deflate_bytes(d, 6, feed=feed, window=window)
        check_stream(oracle, z, d)
        assert z == whole          # input is batched internally: the split does not change the result


def test_sync_flush_keeps_stream_valid(lib, oracle, corpus):
    """DEFLT_FLUSH ends in the byte aligned marker 00 00 FF FF and the instance goes on
    (src/deflator.c:763-768); the concatenation is one valid stream."""
    a, b = corpus.fill(0, 40000), corpus.fill(0, 30000, offset=40000)
    de = lib.deflator(6)
    try:
        za = de.run(a, flush=api.DEFLT_FLUSH)
        assert za[-4:] == b"\x00\x00\xff\xff" and de.state != api.POISON
        dz = zlib.decompressobj(-15)
        assert dz.decompress(za) == a and not dz.eof
        zb = de.run(b, flush=api.DEFLT_END)
        assert de.state == api.POISON
    finally:
        de.close()
    check_stream(oracle, za + zb, a + b)


def test_status_codes_and_misuse(lib, corpus):
    d = corpus.fill(0, 5000)
    buf = bytearray(64)
    de = lib.deflator(6)
    try:
        de.setsrc(d, len(d))
        de.settgt(buf, len(buf))
        assert de.deflate(api.DEFLT_NOFLUSH) == api.SRCEXHSTD
        # again without new input and without a flush: src/deflator.c:663-688
        assert de.deflate(api.DEFLT_NOFLUSH) == api.ERROR and de.error == api.DEFLT_EINCORRECTUSE
        assert de.state == api.POISON
        de.reset()
        de.setsrc(d, len(d))
        de.settgt(buf, len(buf))
        assert de.deflate(api.DEFLT_END) == api.TGTEXHSTD and de.tgtend() == len(buf)
        assert de.deflate(api.DEFLT_END) == api.ERROR and de.error == api.DEFLT_EINCORRECTUSE
        de.reset()
        # END is latched and cannot be downgraded: src/deflator.c:696-699
        de.setsrc(d, len(d))
        big = bytearray(8192)
        de.settgt(big, len(big))
        assert de.deflate(api.DEFLT_END) == api.OK
        assert de.deflate(api.DEFLT_NOFLUSH) == api.ERROR and de.error == 0      # after the end: ERROR, error untouched
    finally:
        de.close()
    assert not lib.lib.deflator_create(0, 10, None)
    assert not lib.lib.deflator_create(0, -1, None)


def test_reset_reuses_instance(lib, oracle, corpus):
    de = lib.deflator(9)
    try:
        for k in (0, 2, 4):
            d = corpus.fill(k, 20000, offset=k)
            z = de.run(d)
            check_stream(oracle, z, d)
            de.reset()
    finally:
        de.close()


def test_multi_batch_pipeline(lib, oracle, corpus, monkeypatch):
    """Several internal batches (two in flight, output drained in order) with small source pieces
    and target windows: the stream must be the same as with one big window."""
    monkeypatch.setenv("JDB200_BATCH_MIB", "1")
    d = corpus.fill(5, (2 << 20) + 12345, offset=(4 << 20) - (1 << 20))
    whole = lib.deflate_bytes(d, 6)
    check_stream(oracle, whole, d)
    z = lib.deflate_bytes(d, 6, feed=300000, window=77777)
    assert z == whole
    # a sync flush in the middle of a batch, then more input
    de = lib.deflator(6)
    try:
        za = de.run(d[:1500000], flush=api.DEFLT_FLUSH, window=500000)
        zb = de.run(d[1500000:], flush=api.DEFLT_END, window=500000)
    finally:
        de.close()
    check_stream(oracle, za + zb, d)


def test_preset_dictionary(lib, corpus):
    """deflator_setdctnr (reference src/deflator.c:2106-2167): the stream needs the same dictionary
    to decode, uses it (smaller than without), and is read back by zlib and by our inflator."""
    dct = corpus.fill(4, 20000, offset=1)
    for n in (300, 5000, 200000):
        d = corpus.fill(4, n, offset=9)                  # same kind of JSON: shares a lot with the dictionary
        for level in (1, 6, 9):
            de = lib.deflator(level)
            try:
                de.setdctnr(dct)
                z = de.run(d)
            finally:
                de.close()
            assert zlib.decompressobj(-15, zdict=dct).decompress(z) == d, (n, level)
            plain = lib.deflate_bytes(d, level)
            assert len(z) < len(plain) or n > 100000, (n, level, len(z), len(plain))
            s = lib.inflator()
            try:
                s.setdctnr(dct)
                st, err, out, used = s.run(z, n)
                assert (st, err, out, used) == (api.OK, 0, d, len(z))
            finally:
                s.close()
            if n <= 5000:
                with pytest.raises(zlib.error):
                    zlib.decompress(z, -15)              # really depends on the dictionary
    # after the first call it is a usage error; at level 0 it is ignored
    de = lib.deflator(6)
    try:
        de.run(b"abc", flush=api.DEFLT_FLUSH)
        de.setdctnr(dct)
        assert de.state == api.POISON and de.error == api.DEFLT_EINCORRECTUSE
    finally:
        de.close()
    de = lib.deflator(0)
    try:
        de.setdctnr(dct)
        assert zlib.decompress(de.run(b"hello"), -15) == b"hello"
    finally:
        de.close()


def test_batch_api_records(lib, oracle, corpus, monkeypatch):
    """jdb200_deflate_batch: every record becomes a complete stream of its own --
    what the reference produces with deflator_reset + deflator_deflate(DEFLT_END) per record.  Each
    stream must decode through zlib, the oracle and our own batched inflate; sizes stay within 3 % of
    the reference's per-record streams (levels 6, 9), and results carry the reference's status codes."""
    recs = ([corpus.json_record(i)[:1500 + 37 * i] for i in range(40)] +
            [b"", b"x", b"ab" * 9, corpus.fill(0, 40000, offset=5), corpus.fill(3, 20000), b"a" * 70000,
             corpus.fill(2, 16384), corpus.fill(1, 16385), corpus.fill(5, 32768, offset=(4 << 20) - 9000)])
    for fmt in (api.JDB200_ZLIB, api.JDB200_RAW):
        for level in (0, 1, 6, 9):
            outs, res = lib.deflate_batch_bytes(recs, fmt=fmt, level=level)
            for r, z, q in zip(recs, outs, res):
                assert (q.status, q.error, q.zerror) == (api.OK, 0, 0), (fmt, level, len(r))
                assert (q.srcused, q.tgtused) == (len(r), len(z))
                if fmt == api.JDB200_ZLIB:
                    assert zlib.decompress(z) == r
                    assert q.checksum == zlib.adler32(r) and (z[0] << 8 | z[1]) % 31 == 0
                    raw = z[2:-4]
                else:
                    assert zlib.decompress(z, -15) == r
                    raw = z
                st = oracle.inflate(raw + b"\x99", len(r) + 1)
                assert (st[0], st[1], st[2], st[3]) == (api.OK, 0, r, len(raw)), (fmt, level, len(r))
            if fmt == api.JDB200_RAW and level in (6, 9):
                ours = sum(len(z) for z in outs)
                ref = sum(len(oracle.deflate(r, level)) for r in recs)
                assert ours <= ref * 1.03, (level, ours, ref)
            # and back through the batched inflate
            back, bres = lib.inflate_batch_bytes(outs, [len(r) for r in recs], fmt=fmt)
            assert back == recs and all(b.status == api.OK and b.zerror == 0 for b in bres)
    # target ranges that are too small: DEFLT_TGTEXHSTD for those records only
    caps = [len(r) + len(r) // 64 + 80 for r in recs]
    caps[3] = 10
    caps[43] = 100
    outs, res = lib.deflate_batch_bytes(recs, caps=caps, fmt=api.JDB200_ZLIB, level=6)
    for i, (r, z, q) in enumerate(zip(recs, outs, res)):
        if i in (3, 43):
            assert (q.status, q.srcused, q.tgtused) == (api.TGTEXHSTD, 0, 0)
        else:
            assert q.status == api.OK and zlib.decompress(z) == r
    # several groups (slot budget of 1 MiB) and a forced slot size smaller than most records
    monkeypatch.setenv("JDB200_BATCH_MIB", "1")
    many = [corpus.fill(4, 3000 + 11 * i, offset=977 * i) for i in range(150)] + [corpus.fill(0, 3 << 20)]
    outs, res = lib.deflate_batch_bytes(many, fmt=api.JDB200_RAW, level=6)
    assert all(q.status == api.OK for q in res) and [zlib.decompress(z, -15) for z in outs] == many
    monkeypatch.setenv("JDB200_RECORD_CHUNK_KIB", "16")
    outs2, res = lib.deflate_batch_bytes(many[:20] + [many[-1][:200000]], fmt=api.JDB200_ZLIB, level=1)
    assert [zlib.decompress(z) for z in outs2] == many[:20] + [many[-1][:200000]]
    # argument errors
    assert lib.lib.jdb200_deflate_batch(None, None, None, None, 3, api.JDB200_RAW, 6) != 0
    assert lib.lib.jdb200_deflate_batch(b"x", b"y", b"z" * 32, b"w" * 32, 1, api.JDB200_RAW, 10) != 0
    assert lib.lib.jdb200_deflate_batch(None, None, None, None, 0, api.JDB200_RAW, 6) == 0


def test_sparse_matches_in_record_structured_data(lib, oracle):
    """lz_kernel samples one position in 16 to tell segments with nothing to find (emitted as literals at
    once) from the rest.  Records of 16 bytes whose only repeating field sits at a fixed offset must not
    alias with that sample: 12 random bytes + a 4-byte field that repeats from record to record."""
    import numpy as np
    rs = np.random.RandomState(11)
    for field_at in (0, 5, 12):
        rec = rs.randint(0, 256, (8192, 16), dtype=np.uint8)
        rec[:, field_at:field_at + 4] = np.frombuffer(b"\xde\xad\xbe\xef", np.uint8)
        d = rec.tobytes()
        z = lib.deflate_bytes(d, 6)
        check_stream(oracle, z, d)
        ref = len(oracle.deflate(d, 6))
        assert ref < 0.97 * len(d)                       # the field is worth compressing
        assert len(z) <= RATIO_TOLERANCE * ref + 16, (field_at, len(z), ref)


def test_preset_dictionary_stream_larger_than_a_chunk_against_the_reference(lib, ref, corpus):
    """DESIGN.md deviation 10: the dictionary is history for the first chunk only.  The reference's sliding window
    carries it 32 KiB far, so on a stream of several chunks both encoders get the same benefit: the sizes stay within
    the 3 % bar, and the reference's inflator (given the dictionary) reads our stream back."""
    dct = corpus.fill(4, 30000, offset=1)
    d = corpus.fill(4, 3 * 512 * 1024 + 12345, offset=9)
    sizes = {}
    for name, l in (("ours", lib), ("ref", ref)):
        de = l.deflator(6)
        try:
            de.setdctnr(dct)
            z = de.run(d)
        finally:
            de.close()
        assert zlib.decompressobj(-15, zdict=dct).decompress(z) == d, name
        sizes[name] = len(z)
        if name == "ours":
            s = ref.inflator()
            try:
                s.setdctnr(dct)
                st, err, out, used = s.run(z, len(d))
                assert (st, err, used) == (api.OK, 0, len(z)) and out == d
            finally:
                s.close()
    assert sizes["ours"] <= RATIO_TOLERANCE * sizes["ref"], sizes


def test_chain_links_equal_the_exact_model(emu, corpus):
    """chain_kernel + chain_fix_kernel against a plain restatement of what a link is (the reference's
    hash insert, src/deflator.c:2402-2428: every position points at the previous position with the
    same hash of its 4 bytes, inside the window and inside its chunk): with and without the head
    tables that replace the warm-up replay, ranges of 64 and 128 KiB, a ragged tail.  Pins two things
    that went wrong once: the start value of the head table reading as a link to position 32768, and
    the links across range boundaries."""
    import ctypes as C
    import numpy as np
    lib = emu.lib
    lib.jdb_dev_alloc.restype = C.c_void_p
    lib.jdb_dev_alloc.argtypes = [C.c_size_t]
    lib.jdb_lz_chain.restype = C.c_int
    lib.jdb_lz_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.jdb_lz_chain_heads_bytes.restype = C.c_size_t
    lib.jdb_lz_chain_heads_bytes.argtypes = [C.c_uint64, C.c_uint32]

    def links(data, chunk, rng, use_heads):
        n = len(data)
        npad = (n + 8191) // 8192 * 8192 + 64
        din = lib.jdb_dev_alloc(npad)
        C.memset(din, 0, npad)
        C.memmove(din, data, n)
        prev = lib.jdb_dev_alloc(npad * 2)
        C.memset(prev, 0, npad * 2)
        heads = lib.jdb_dev_alloc(lib.jdb_lz_chain_heads_bytes(n, chunk)) if use_heads else None
        assert lib.jdb_lz_chain(din, n, chunk, rng, None, prev, heads, None) == 0
        return np.frombuffer((C.c_uint16 * n).from_address(prev), np.uint16).astype(np.int64)

    def model(data, chunk):
        d = np.frombuffer(data, np.uint8).astype(np.uint64)
        n = len(d)
        be = (d[:-3] << 24) | (d[1:-2] << 16) | (d[2:-1] << 8) | d[3:]
        h = ((be * 0x1e35a7bd) & 0xffffffff) >> (32 - 14)
        want = np.zeros(n, np.int64)
        last = {}
        for p in range(n - 3):
            if p % chunk == 0:
                last = {}
            if p + 3 >= min((p // chunk + 1) * chunk, n):
                continue
            q = last.get(int(h[p]))
            if q is not None and p - q < 32768:
                want[p] = p - q
            last[int(h[p])] = p
        return want

    for kind, n in ((0, 262144 + 131072 + 5000), (2, 200000), (4, 70000)):
        data = corpus.fill(kind, n, offset=99)
        for chunk, rng in ((262144, 65536), (524288, 131072)):
            want = model(data, chunk)
            for use_heads in (False, True):
                got = links(data, chunk, rng, use_heads)
                bad = np.nonzero(got != want)[0]
                assert len(bad) == 0, (kind, n, chunk, rng, use_heads, bad[:5], got[bad[:5]], want[bad[:5]])
