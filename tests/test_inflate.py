"""Raw DEFLATE decode parity (reference inflator_inflate, src/inflator.c:764-903).
Oracle: oracle/jd_oracle.c (pinned in test_oracle.py) and zlib; decoded bytes bit exact,
status / error codes equal, input accounting exact."""
import base64
import random
import zlib

import pytest

from jdeflate_b200 import api
from support import KIND_NAMES, zlib_raw
from test_oracle import RFC_DEVIATIONS


def test_known_answer_streams(lib, golden):
    for k in golden["inflate_kat"]:
        s = base64.b64decode(k["stream"])
        if not s:
            continue
        st, err, out, used = lib.inflate_bytes(s, k["cap"], final=bool(k["final"]))
        want = RFC_DEVIATIONS.get(k["name"], (k["status"], k["error"]))
        assert (st, err) == want, k["name"]
        if k["name"] in RFC_DEVIATIONS:
            continue
        assert len(out) == k["out_len"] and zlib.crc32(out) == k["out_crc"], k["name"]
        if st == api.OK:
            assert used == k["exact_consumed"], k["name"]


@pytest.mark.parametrize("kind", range(5))
def test_third_party_and_reference_streams(lib, oracle, corpus, kind):
    for n in (1, 100, 4096, 70000, 200000):
        d = corpus.fill(kind, n, offset=17 * n)
        streams = [zlib_raw(d, lvl) for lvl in (1, 6, 9)] + [oracle.deflate(d, lvl) for lvl in (0, 1, 6)]
        for z in streams:
            st, err, out, used = lib.inflate_bytes(z + b"\x55\xaa", n + 1)
            want = oracle.inflate(z + b"\x55\xaa", n + 1)
            assert (st, err, used) == (want[0], want[1], want[3]) == (api.OK, 0, len(z)), (KIND_NAMES[kind], n)
            assert out == d


def test_streaming_small_windows(lib, corpus):
    """Resumability: any split of source and target windows gives the same bytes
    (reference substate machinery, src/inflator.c:105-114)."""
    big = corpus.fill(4, 50000, offset=5)
    small = big[:3000]
    for d, feed, window in ((small, 1, 3001), (small, 7, 13), (small, None, 1), (big, 1000, 300), (big, 333, 40000)):
        z = zlib_raw(d, 6)
        st, err, out, used = lib.inflate_bytes(z, len(d), window=window, feed=feed)
        assert (st, err) == (api.OK, 0), (feed, window)
        assert out == d and used == len(z)


def test_truncation_and_final_flag(lib, corpus):
    d = corpus.fill(0, 30000)
    z = zlib_raw(d, 6)
    half = z[: len(z) // 2]
    st, err, out, _ = lib.inflate_bytes(half, len(d), final=False)
    assert (st, err) == (api.SRCEXHSTD, 0) and d.startswith(out) and len(out) > 0
    st, err, out, _ = lib.inflate_bytes(half, len(d), final=True)
    assert (st, err) == (api.ERROR, api.INFLT_EINPUTEND)
    st, err, out, _ = lib.inflate_bytes(z, 1000)
    assert st == api.TGTEXHSTD and out == d[:1000]


def test_corrupted_streams_match_oracle(lib, oracle, corpus):
    rnd = random.Random(99)
    d = corpus.fill(4, 5000, offset=3)
    z = zlib_raw(d, 6)
    for _ in range(120):
        m = bytearray(z)
        for _ in range(rnd.randint(1, 3)):
            m[rnd.randrange(len(m))] ^= 1 << rnd.randrange(8)
        want = oracle.inflate(bytes(m), 6000)
        got = lib.inflate_bytes(bytes(m), 6000)
        assert (got[0], got[1]) == (want[0], want[1])
        if want[0] == api.OK:
            assert got[2] == want[2] and got[3] == want[3]
        else:
            assert got[2] == want[2][: len(got[2])] or want[2] == got[2][: len(want[2])]


def test_sticky_error_and_misuse(lib, corpus):
    s = lib.inflator()
    try:
        bad = bytes([0x07, 0, 0])                      # BFINAL=1, BTYPE=3
        buf = bytearray(16)
        s.setsrc(bad, len(bad))
        s.settgt(buf, len(buf))
        assert s.inflate(1) == api.ERROR and s.error == api.INFLT_EBADBLOCK and s.state == api.POISON
        assert s.inflate(1) == api.ERROR and s.error == api.INFLT_EBADBLOCK
        s.reset()
        assert s.state == 0 and s.error == 0
        z = zlib_raw(b"hello world, hello world", 6)
        s.setsrc(z, len(z))
        s.settgt(buf, 4)
        assert s.inflate(1) == api.TGTEXHSTD
        # asking again without a fresh target window is a usage error (src/inflator.c:729-762)
        assert s.inflate(1) == api.ERROR and s.error == api.INFLT_EINCORRECTUSE
    finally:
        s.close()


def test_preset_dictionary(lib, corpus):
    dct = corpus.fill(0, 20000, offset=7)
    d = dct[5000:9000] + corpus.fill(0, 3000, offset=999999)
    co = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=dct)
    z = co.compress(d) + co.flush()
    s = lib.inflator()
    try:
        s.setdctnr(dct)
        st, err, out, used = s.run(z, len(d))
        assert (st, err, out, used) == (api.OK, 0, d, len(z))
    finally:
        s.close()


def test_batch_api(lib, oracle, corpus):
    """jdb200_inflate_batch: per record results equal a fresh inflator fed the record with final = 1."""
    recs = [corpus.json_record(i) for i in range(40)] + [b"", b"x"]
    streams = [zlib.compress(r, 6) for r in recs]
    outs, res = lib.inflate_batch_bytes(streams, [len(r) for r in recs], fmt=api.JDB200_ZLIB)
    for r, z, o, q in zip(recs, streams, outs, res):
        assert (q.status, q.error, q.zerror) == (api.OK, 0, 0)
        assert o == r and q.srcused == len(z) and q.tgtused == len(r) and q.checksum == zlib.adler32(r)
    # raw format, with damaged, truncated and over-long members in the same batch
    raw = [zlib_raw(r, 6) for r in recs[:8]]
    bad = bytearray(raw[3]); bad[len(bad) // 2] ^= 0x40
    raw[3] = bytes(bad)
    raw[5] = raw[5][: len(raw[5]) // 2]
    caps = [len(r) for r in recs[:8]]
    caps[6] = 100                                        # target too small
    outs, res = lib.inflate_batch_bytes(raw, caps, fmt=api.JDB200_RAW)
    for i in range(8):
        want = oracle.inflate(raw[i], caps[i], final=True)
        assert (res[i].status, res[i].error) == (want[0], want[1]), i
        if want[0] == api.OK:
            assert outs[i] == want[2]
    assert res[5].status == api.ERROR and res[5].error == api.INFLT_EINPUTEND
    assert res[6].status == api.TGTEXHSTD and outs[6] == recs[6][:100]
    # checksum mismatch in a zlib member
    z = bytearray(streams[0]); z[-1] ^= 1
    _, res = lib.inflate_batch_bytes([bytes(z)], [len(recs[0])], fmt=api.JDB200_ZLIB)
    assert res[0].status == api.OK and res[0].zerror == api.ZSTRM_ECHECKSUM


def _flushed_stream(data, piece, mode):
    """Raw deflate by zlib with a flush of `mode` after every `piece` bytes."""
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    out = bytearray()
    for off in range(0, len(data), piece):
        out += co.compress(data[off:off + piece])
        out += co.flush(mode)
    out += co.flush(zlib.Z_FINISH)
    return bytes(out)


def test_chunk_parallel_decode_of_marker_cut_streams(lib, oracle, corpus, monkeypatch):
    """Streams cut into independent chunks by sync markers (ours; zlib's Z_FULL_FLUSH) are decoded
    chunk-parallel inside the plain inflator; the result -- bytes, status,
    consumed input -- must be what the sequential decoder gives, for every shape of target window,
    and streams that merely LOOK chunked (Z_SYNC_FLUSH: matches cross the markers; marker bytes
    inside stored data) must fall back cleanly."""
    monkeypatch.setenv("JDB200_CHUNK_KIB", "64")
    d = corpus.fill(5, 1200000, offset=(4 << 20) - 500000)          # TEXT then BINARY
    ours = lib.deflate_bytes(d, 6)
    assert ours.count(b"\x00\x00\xff\xff") >= 18
    full = _flushed_stream(d, 50000, zlib.Z_FULL_FLUSH)
    sync = _flushed_stream(d, 50000, zlib.Z_SYNC_FLUSH)
    noise = b"\x00\x00\xff\xff" * 2000 + corpus.fill(3, 300000)       # marker bytes as payload
    stored = zlib_raw(noise, 0)
    for name, z, want in (("ours", ours, d), ("full_flush", full, d), ("sync_flush", sync, d), ("stored_noise", stored, noise)):
        assert len(z) >= 256 << 10, name
        for window in (None, 1 << 20, 100000, 65536 + 7):
            st, err, out, used = lib.inflate_bytes(z + b"tail", len(want), window=window)
            assert (st, err, used) == (api.OK, 0, len(z)), (name, window)
            assert out == want, (name, window)
        # not final, input ends in the middle of a chunk: everything decodable comes out, then SRCEXHSTD
        cut = len(z) * 2 // 3
        st, err, out, used = lib.inflate_bytes(z[:cut], len(want), final=False)
        ost = oracle.inflate(z[:cut], len(want), final=False)
        assert (st, err) == (api.SRCEXHSTD, 0) and out == ost[2], name
    # damage inside a chunk of our own stream: the error of the sequential decoder
    bad = bytearray(ours)
    bad[len(bad) // 2] ^= 0x5a
    got = lib.inflate_bytes(bytes(bad), len(d))
    exp = oracle.inflate(bytes(bad), len(d))
    assert (got[0], got[1]) == (exp[0], exp[1])
    if exp[0] == api.OK:
        assert got[2] == exp[2]


def test_parallel_path_not_final_input(lib, corpus):
    """Regression (round-1 advice): after a chunk-parallel step the inflator must not ask for more
    input while the queue still holds the complete rest of the stream.  final = 0 throughout, as
    zstrm calls it; the reference returns INFLT_OK for the same calls."""
    d = corpus.fill(0, 4392960, offset=1 << 20)
    for wbits in (-15, 15, 31):
        co = zlib.compressobj(6, zlib.DEFLATED, wbits)
        z = co.compress(d[:307200]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(d[307200:]) + co.flush()
        assert len(z) >= 1 << 20
        if wbits == -15:
            st, err, out, used = lib.inflate_bytes(z, len(d), final=False)
            assert (st, err, used) == (api.OK, 0, len(z))
            assert out == d
        else:
            zs = lib.zstrm(api.ZSTRM_INFLATE | (api.ZSTRM_ZLIB if wbits == 15 else api.ZSTRM_GZIP))
            zs.setsource(z)
            out = zs.inflate(len(d) + 16)
            assert zs.error == 0, wbits
            assert out == d, wbits
            zs.close()


def test_parallel_path_host_target_larger_than_staging(lib, corpus):
    """Regression (round-1 advice): a chain of chunks longer than what one pass can hand to a host
    target (64 MiB) is delivered over several passes of the same call -- final = 0, all input given."""
    n = 66 << 20
    d = corpus.fill(3, 1 << 20) * 66
    z = lib.deflate_bytes(d, 0)
    st, err, out, used = lib.inflate_bytes(z, n, final=False)
    assert (st, err, used) == (api.OK, 0, len(z))
    assert len(out) == n and zlib.crc32(out) == zlib.crc32(d)
    # the same stream through small target windows: the decoded chunks wait in their slots
    st, err, out, used = lib.inflate_bytes(z[: 9 << 20] , 9 << 20, window=(1 << 20) + 13, final=False)
    assert st == api.SRCEXHSTD and err == 0
    assert out == d[: len(out)] and len(out) >= 8 << 20


def test_parallel_path_small_source_windows(lib, corpus, monkeypatch):
    """A streaming caller that feeds our own stream in windows smaller than a parallel step wants:
    bytes and status as the sequential decoder, whichever path ends up decoding."""
    monkeypatch.setenv("JDB200_CHUNK_KIB", "64")
    d = corpus.fill(0, 6 << 20, offset=3 << 20)
    z = lib.deflate_bytes(d, 6)
    for feed in (1 << 20, 300000, 65536):
        st, err, out, used = lib.inflate_bytes(z, len(d), feed=feed, final=False)
        assert (st, err, used) == (api.OK, 0, len(z)), feed
        assert out == d, feed


# ---- the lane-parallel rounds, the output ring and the word copies (round 2) ----------------------

def test_lane_parallel_rounds_shapes(lib, oracle):
    """Inputs that push the lane-parallel symbol decode to its edges: one- and two-bit codes (a lane's
    token slots run out long before its subsequence ends), maximal matches, literal-only blocks with the
    fixed code, a long run (distance 1, overlapping copies), blocks that end inside a round."""
    import numpy as np
    rs = np.random.RandomState(3)
    cases = {
        "zeros": bytes(300000),
        "two_symbols": bytes(rs.randint(0, 2, 200000, dtype=np.uint8) * 65),
        "four_symbols": bytes(rs.randint(0, 4, 200000, dtype=np.uint8) + 97),
        "random_literals": rs.randint(0, 256, 150000, dtype=np.uint8).tobytes(),
        "period_258": (bytes(range(256)) + b"ab") * 600,
        "long_run_then_noise": b"\x07" * 70000 + rs.randint(0, 256, 3000, dtype=np.uint8).tobytes() + b"\x07" * 70000,
    }
    for name, d in cases.items():
        for lvl, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (9, zlib.Z_RLE)):
            c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strategy)
            z = c.compress(d) + c.flush()
            st, err, out, used = lib.inflate_bytes(z, len(d) + 1)
            assert (st, err, used) == (api.OK, 0, len(z)), (name, lvl, strategy)
            assert out == d, (name, lvl, strategy)
    # many short blocks: Z_FULL_FLUSH every few hundred bytes puts block ends (and stored markers) inside rounds
    d = cases["four_symbols"][:60000]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    z = b"".join(c.compress(d[i:i + 700]) + c.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(d), 700)) + c.flush()
    st, err, out, used = lib.inflate_bytes(z, len(d))
    assert (st, err, used) == (api.OK, 0, len(z)) and out == d


def test_target_windows_cut_lane_parallel_rounds(lib, corpus):
    """The target room ends inside a round: the chain is cut, the step-by-step decoder finishes the window,
    a match cut by the window end is continued by the next call."""
    d = corpus.fill(4, 40000, offset=11) + bytes(5000) + corpus.fill(0, 20000, offset=3)
    z = zlib_raw(d, 6)
    for window in (255, 258, 259, 1000, 4095, 4096, 4097, 33000):
        st, err, out, used = lib.inflate_bytes(z, len(d), window=window)
        assert (st, err, used) == (api.OK, 0, len(z)), window
        assert out == d, window
    # the smallest windows on a shorter stream (one launch per window)
    ds = d[36000:48000]
    zs = zlib_raw(ds, 6)
    for window in (1, 2, 3, 5, 17):
        st, err, out, used = lib.inflate_bytes(zs, len(ds), window=window)
        assert (st, err, used) == (api.OK, 0, len(zs)), window
        assert out == ds, window
    # exact fit, one byte short, and a target that ends inside a long match
    for cap in (len(d), len(d) - 1, 40000 + 2500):
        st, err, out, used = lib.inflate_bytes(z, cap)
        assert st == (api.OK if cap == len(d) else api.TGTEXHSTD) and out == d[:cap], cap


def test_far_offset_inside_a_round_matches_oracle(lib, oracle):
    """A distance that reaches beyond the bytes produced so far, deep inside the input (not at its start): the
    round drops the lane that holds it and the step-by-step decoder reports the reference's error and output."""
    from support import BitWriter
    for lits in (40, 300, 2000):
        w = BitWriter()
        w.put(1, 1); w.put(1, 2)                        # final, fixed code
        def lit(b):
            if b < 144: w.huff(0x30 + b, 8)
            else: w.huff(0x190 + b - 144, 9)
        for i in range(lits):
            lit((i * 7) & 0x7f)
        # length 3 (symbol 257, 7 bits 0000001), distance code 29 + 13 extra bits all ones = 32768
        w.huff(1, 7); w.huff(29, 5); w.put(0x1fff, 13)
        for i in range(64):
            lit(65)
        w.huff(0, 7)                                    # end of block
        z = w.bytes()
        got = lib.inflate_bytes(z, lits + 200)
        want = oracle.inflate(z, lits + 200)
        assert got[:2] == (want[0], want[1]) == (api.ERROR, api.INFLT_EFAROFFSET), lits
        assert got[2] == want[2] and len(got[2]) == lits


def test_batch_records_at_every_alignment(lib, corpus):
    """Output ring words line up with target words: records whose target offsets and lengths hit every
    alignment, stored and compressed, through the batch call."""
    import numpy as np
    rs = np.random.RandomState(5)
    recs, streams = [], []
    for i in range(48):
        n = 1 + (i * 37) % 300 + (i % 5) * 1000
        r = corpus.fill(i % 5, n, offset=i * 1000) if i % 3 else rs.randint(0, 256, n, dtype=np.uint8).tobytes()
        recs.append(r)
        streams.append(zlib_raw(r, 0 if i % 4 == 0 else 6))
    outs, res = lib.inflate_batch_bytes(streams, [len(r) for r in recs], fmt=api.JDB200_RAW)
    for i, (r, o, q) in enumerate(zip(recs, outs, res)):
        assert (q.status, q.error) == (api.OK, 0) and o == r, i


# ---- one stream on a whole CTA (inflate_wide_kernel: the sequential path of inflator_inflate) ----

def _mixed_stream(corpus, n, level=6):
    """Text, binary, logs, random and a run of zeros in one stream: dynamic, fixed and stored blocks,
    short and 258-byte matches, distances up to the window."""
    import numpy as np
    parts, k = [], 0
    while sum(map(len, parts)) < n:
        kind = k % 6
        m = 30000 + 7919 * (k % 7)
        if kind == 5:
            parts.append(bytes(m // 3) if k % 2 else np.random.RandomState(k).randint(0, 256, m // 4, dtype=np.uint8).tobytes())
        else:
            parts.append(corpus.fill(kind % 5, m, offset=k * 100003))
        k += 1
    d = b"".join(parts)[:n]
    return d, zlib_raw(d, level)


def test_wide_rounds_long_stream(lib, oracle, corpus):
    """A third-party stream long enough for full rounds of 512 lanes, in one call and in windows that
    cut rounds short at every kind of place (target room, source end), against zlib's bytes and the
    oracle's status / accounting."""
    d, z = _mixed_stream(corpus, 1_300_000)
    st, err, out, used = lib.inflate_bytes(z + b"\x01\x02\x03", len(d) + 5)
    assert (st, err, used) == (api.OK, 0, len(z)) and out == d
    want = oracle.inflate(z + b"\x01\x02\x03", len(d) + 5)
    assert (want[0], want[1], want[3]) == (st, err, used)
    for feed, window in ((None, 70001), (33333, 65536), (4099, 1 << 20), (100000, 9973)):
        st, err, out, used = lib.inflate_bytes(z, len(d), window=window, feed=feed)
        assert (st, err, used) == (api.OK, 0, len(z)), (feed, window)
        assert out == d, (feed, window)
    # the target ends inside the stream: exactly cap bytes, TGTEXHSTD
    for cap in (1, 4097, 65537, 700001):
        st, err, out, _ = lib.inflate_bytes(z, cap)
        assert st == api.TGTEXHSTD and out == d[:cap], cap


def test_wide_rounds_levels_and_kinds(lib, corpus):
    for kind in range(5):
        d = corpus.fill(kind, 400_000, offset=kind * 31)
        for lvl in (1, 9):
            z = zlib_raw(d, lvl)
            st, err, out, used = lib.inflate_bytes(z, len(d))
            assert (st, err, used) == (api.OK, 0, len(z)) and out == d, (kind, lvl)


def test_wide_rounds_history_and_dictionary(lib, corpus):
    """Distances that reach into earlier calls' output (the history ring of the device state) and
    into a preset dictionary: the wide rounds mirror the last 32 KiB in shared memory."""
    dct = corpus.fill(0, 32768, offset=11)
    d = dct[1000:30000] + corpus.fill(0, 300_000, offset=424242) + dct[:20000]
    co = zlib.compressobj(9, zlib.DEFLATED, -15, zdict=dct)
    z = co.compress(d) + co.flush()
    for feed, window in ((None, None), (5000, 20011), (77777, 300)):
        s = lib.inflator()
        try:
            s.setdctnr(dct)
            st, err, out, used = s.run(z, len(d), window=window, feed=feed)
            assert (st, err, used) == (api.OK, 0, len(z)) and out == d, (feed, window)
        finally:
            s.close()


def test_wide_rounds_corruption_parity(lib, oracle, corpus):
    """Damage deep inside a long stream: status, error and the bytes in front of it as the oracle has them."""
    rnd = random.Random(4242)
    d, z = _mixed_stream(corpus, 400_000)
    for _ in range(25):
        m = bytearray(z)
        for _ in range(rnd.randint(1, 2)):
            m[rnd.randrange(len(m) // 3, len(m))] ^= 1 << rnd.randrange(8)
        want = oracle.inflate(bytes(m), len(d) + 100)
        got = lib.inflate_bytes(bytes(m), len(d) + 100)
        assert (got[0], got[1]) == (want[0], want[1])
        if want[0] == api.OK:
            assert got[2] == want[2] and got[3] == want[3]
        else:
            assert got[2] == want[2][: len(got[2])] or want[2] == got[2][: len(want[2])]


@pytest.mark.gpu
def test_wide_and_narrow_decoders_agree_on_device_targets(jd, corpus):
    """Device buffers at every target alignment: the CTA-wide decoder writes aligned words from its entry
    array, the one-warp decoder from its ring -- same bytes, same accounting."""
    import os
    import torch
    d, z = _mixed_stream(corpus, 3_000_000)
    src = torch.frombuffer(bytearray(z), dtype=torch.uint8).cuda()
    want = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    for narrow in ("0", "1"):
        os.environ["JDB200_INFLATE_NARROW"] = narrow
        try:
            for off in (0, 1, 2, 3):
                back = torch.zeros(len(d) + 8, dtype=torch.uint8, device="cuda")
                s = jd.inflator()
                try:
                    s.setsrc(src.data_ptr(), len(z))
                    s.settgt(back.data_ptr() + off, len(d))
                    assert s.inflate(1) == api.OK and s.tgtend() == len(d) and s.srcend() == len(z)
                finally:
                    s.close()
                assert torch.equal(back[off:off + len(d)], want), (narrow, off)
                assert int(back[:off].sum()) == 0 and int(back[off + len(d):].sum()) == 0
        finally:
            os.environ.pop("JDB200_INFLATE_NARROW", None)


def test_small_batches_take_a_thread_block_per_stream(lib, oracle, corpus, monkeypatch):
    """jdb200_inflate_batch with fewer streams than SMs runs inflate_wide_kernel (one CTA per stream), with the
    zlib framing of the batch kernel: good, damaged, truncated and over-long members, a bad header, a preset
    dictionary flag, a target that is too small -- every result field as the one-warp kernel reports it, and
    as the oracle decides for the raw streams."""
    monkeypatch.setenv("JDB_EMU_SMS", "148")        # the emulator's device has 4 SMs by default; a B200 has 148
    d_long, _ = _mixed_stream(corpus, 700_000)
    recs = [b"", b"x", corpus.fill(0, 5000, offset=1), corpus.fill(2, 70000, offset=2), d_long, corpus.json_record(7)]
    good = [zlib.compress(r, 6) for r in recs]
    bad_sum = bytearray(good[3]); bad_sum[-1] ^= 1
    bad_data = bytearray(good[4]); bad_data[len(bad_data) // 2] ^= 0x10
    bad_hdr = bytearray(good[2]); bad_hdr[0] = 0x79
    fdict = bytearray(good[2]); fdict[1] |= 0x20
    streams = good + [bytes(bad_sum), bytes(bad_data), good[4][: len(good[4]) // 3], bytes(bad_hdr), bytes(fdict), good[3], b"\x78"]
    want_out = recs + [recs[3], None, None, None, None, None, None]
    caps = [len(r) for r in recs] + [len(recs[3]), len(recs[4]), len(recs[4]), 5000, 5000, 1000, 10]
    outs, res = lib.inflate_batch_bytes(streams, caps, fmt=api.JDB200_ZLIB)
    for i in range(len(recs)):
        q = res[i]
        assert (q.status, q.error, q.zerror) == (api.OK, 0, 0), i
        assert outs[i] == recs[i] and q.srcused == len(streams[i]) and q.checksum == zlib.adler32(recs[i]), i
    k = len(recs)
    assert (res[k].status, res[k].zerror) == (api.OK, 4) and outs[k] == recs[3]            # checksum mismatch
    assert res[k + 2].status == api.ERROR and res[k + 2].error == api.INFLT_EINPUTEND     # truncated
    assert res[k + 3].zerror == 3 and res[k + 4].zerror == 6 and res[k + 6].zerror == 3   # header, FDICT, 1 byte
    assert res[k + 5].status == api.TGTEXHSTD and outs[k + 5] == recs[3][:1000]
    # the same batch through the one-warp kernel (a device with a single SM has fewer SMs than streams)
    monkeypatch.setenv("JDB_EMU_SMS", "1")
    outs1, res1 = lib.inflate_batch_bytes(streams, caps, fmt=api.JDB200_ZLIB)
    fields = ("status", "error", "zerror", "checksum", "srcused", "tgtused")
    for i, (a, b) in enumerate(zip(res, res1)):
        assert [getattr(a, f) for f in fields] == [getattr(b, f) for f in fields], i
        assert outs[i] == outs1[i], i
    # raw streams against the oracle
    monkeypatch.setenv("JDB_EMU_SMS", "148")
    raw = [z[2:-4] for z in good] + [bytes(bad_data)[2:-4], good[4][2: len(good[4]) // 3]]
    rcaps = [len(r) for r in recs] + [len(recs[4]), len(recs[4])]
    outs, res = lib.inflate_batch_bytes(raw, rcaps, fmt=api.JDB200_RAW)
    for i, (z, cap) in enumerate(zip(raw, rcaps)):
        want = oracle.inflate(z, cap, final=True)
        assert (res[i].status, res[i].error) == (want[0], want[1]), i
        if want[0] == api.OK:
            assert outs[i] == want[2] and res[i].srcused == want[3], i
