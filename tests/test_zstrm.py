"""gzip / zlib / raw container layer (reference zstrm, src/zstrm.c:80-1313).

Oracle: Python's zlib / gzip modules (zlib 1.3) for the container formats and checksums (the
reference's own zlib header and Adler-32 are defective), the compiled
reference for cross-decoding where it is present."""
import gzip
import zlib

import pytest

from jdeflate_b200 import api

TYPES = [(api.ZSTRM_GZIP, "gzip"), (api.ZSTRM_ZLIB, "zlib"), (api.ZSTRM_DFLT, "raw")]


def compress(lib, data, typ, level=6, piece=30000, flags=0):
    out = bytearray()
    calls = []
    z = lib.zstrm(api.ZSTRM_DEFLATE | typ | flags, level)
    try:
        z.settargetfn(lambda b: (out.extend(b), calls.append(len(b)), len(b))[2])
        for pos in range(0, len(data), piece):
            chunk = data[pos:pos + piece]
            assert z.deflate(chunk) == len(chunk)
        z.flush(1)
        assert (z.error, z.state) == (0, 4)
        return bytes(out), (z.s.crc, z.s.adler, z.s.total)
    finally:
        z.close()


def decompress(lib, comp, read=7777, via_callback=None, flags=0):
    z = lib.zstrm(api.ZSTRM_INFLATE | flags)
    try:
        if via_callback:
            st = {"p": 0}

            def rd(size):
                k = min(size, via_callback, len(comp) - st["p"])
                b = comp[st["p"]: st["p"] + k]
                st["p"] += k
                return b
            z.setsourcefn(rd)
        else:
            z.setsource(comp)
        got = bytearray()
        while z.state != 4:
            b = z.inflate(read)
            got += b
            if not b:
                break
        return bytes(got), z.error, (z.s.stype, z.s.crc, z.s.adler, z.s.total, z.s.usedinput)
    finally:
        z.close()


def third_party_decode(comp, name):
    if name == "gzip":
        return gzip.decompress(comp)
    if name == "zlib":
        return zlib.decompress(comp)
    return zlib.decompress(comp, -15)


@pytest.mark.parametrize("typ,name", TYPES)
def test_roundtrip_all_containers(lib, corpus, typ, name):
    for kind, n in ((0, 100000), (2, 200000), (3, 5000), (4, 1), (0, 0)):
        d = corpus.fill(kind, n, offset=n) if n else b""
        comp, (crc, adler, total) = compress(lib, d, typ)
        assert third_party_decode(comp, name) == d
        assert total == n
        if name == "gzip":
            assert comp[:10] == bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0])
            assert crc == zlib.crc32(d)
            assert comp[-8:] == zlib.crc32(d).to_bytes(4, "little") + (n & 0xffffffff).to_bytes(4, "little")
        if name == "zlib":
            # valid FCHECK (the reference writes 78 1F, src/zstrm.c:1038; DESIGN.md deviation 1)
            assert comp[0] == 0x78 and ((comp[0] << 8) | comp[1]) % 31 == 0
            assert adler == zlib.adler32(d) and comp[-4:] == zlib.adler32(d).to_bytes(4, "big")
        for kw in ({}, {"via_callback": 5000}, {"read": 1 << 20}, {"read": 1 << 20, "via_callback": 1 << 20}):
            back, err, (stype, c2, a2, t2, used) = decompress(lib, comp + b"tail", **kw)
            assert (back, err) == (d, 0), (name, kw)
            assert stype == typ and t2 == n and used == len(comp)
            if name == "gzip":
                assert c2 == zlib.crc32(d)
            if name == "zlib":
                assert a2 == zlib.adler32(d)


def test_third_party_streams(lib, corpus):
    """Streams made by zlib / gzip (the decoder side of BASELINE config 3 and 5)."""
    d = corpus.fill(4, 150000, offset=9)
    for comp, typ in ((zlib.compress(d, 6), api.ZSTRM_ZLIB), (gzip.compress(d, 6), api.ZSTRM_GZIP),
                      (zlib.compress(d, 1), api.ZSTRM_ZLIB)):
        back, err, info = decompress(lib, comp)
        assert (back, err, info[0], info[4]) == (d, 0, typ, len(comp))
    # gzip member with FNAME + FEXTRA + FCOMMENT + FHCRC fields is skipped over (src/zstrm.c:479-503)
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(d) + raw.flush()
    head = bytes([0x1f, 0x8b, 8, 0x02 | 0x04 | 0x08 | 0x10, 1, 2, 3, 4, 0, 3]) + b"\x05\x00extra" + b"name.txt\x00" + b"a comment\x00" + b"\x12\x34"
    comp = head + body + zlib.crc32(d).to_bytes(4, "little") + (len(d) & 0xffffffff).to_bytes(4, "little")
    back, err, info = decompress(lib, comp)
    assert (back, err, info[4]) == (d, 0, len(comp))


def test_trailer_and_format_errors(lib, corpus):
    d = corpus.fill(0, 20000)
    g = bytearray(gzip.compress(d))
    g[-5] ^= 1                                                   # CRC-32
    assert decompress(lib, bytes(g))[1] == api.ZSTRM_ECHECKSUM
    g = bytearray(gzip.compress(d))
    g[-1] ^= 1                                                   # ISIZE
    assert decompress(lib, bytes(g))[1] == api.ZSTRM_EBADDATA
    z = bytearray(zlib.compress(d))
    z[-1] ^= 1                                                   # Adler-32
    assert decompress(lib, bytes(z))[1] == api.ZSTRM_ECHECKSUM
    assert decompress(lib, bytes(z), flags=api.ZSTRM_NOADLER)[:2] == (d, 0)
    # accepted-type mask: a gzip stream offered to a zlib-only reader (src/zstrm.c:598-601)
    assert decompress(lib, gzip.compress(d), flags=api.ZSTRM_ZLIB)[1] == api.ZSTRM_EFORMAT
    # corrupted body
    z = bytearray(zlib.compress(d))
    z[len(z) // 2] ^= 0x10
    assert decompress(lib, bytes(z))[1] in (api.ZSTRM_EDEFLATE, api.ZSTRM_ECHECKSUM)
    # truncated memory source
    assert decompress(lib, zlib.compress(d)[:500])[1] == api.ZSTRM_ESRCEXHSTD
    # truncated callback source: 0 from the callback mid stream (src/zstrm.c:875-878)
    assert decompress(lib, zlib.compress(d)[:500], via_callback=100)[1] == api.ZSTRM_EBADDATA
    assert decompress(lib, b"\x1f\x8c\x08" + bytes(20))[1] == api.ZSTRM_EBADDATA


def test_create_validation_and_misuse(lib):
    L = lib.lib
    assert not L.zstrm_create(0, 6, None)                                            # no mode
    assert not L.zstrm_create(api.ZSTRM_DEFLATE, 6, None)                            # deflate needs one type
    assert not L.zstrm_create(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP | api.ZSTRM_ZLIB, 6, None)
    assert not L.zstrm_create(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, 10, None)
    z = lib.zstrm(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, 6)
    try:
        z.setsource(b"abc")                          # wrong direction
        assert (z.error, z.state) == (api.ZSTRM_EINCORRECTUSE, 4)
        z.reset()
        assert (z.error, z.state) == (0, 0)
        z.settargetfn(lambda b: -1)                  # failing callback
        assert z.deflate(b"x" * 100) == 0 and z.error == api.ZSTRM_EIOERROR
    finally:
        z.close()
    z = lib.zstrm(api.ZSTRM_INFLATE)
    try:
        z.setsource(zlib.compress(b"hello"))
        assert z.s.stype == api.ZSTRM_ZLIB and z.state == 3          # header parsed by setsource
        assert z.inflate(100) == b"hello" and z.state == 4 and z.error == 0
    finally:
        z.close()


def test_zlib_preset_dictionary_inflate(lib, corpus):
    dct = corpus.fill(0, 8000, offset=3)
    d = dct[1000:3000] + corpus.fill(0, 2000, offset=123456)
    co = zlib.compressobj(6, zdict=dct)
    comp = co.compress(d) + co.flush()
    z = lib.zstrm(api.ZSTRM_INFLATE)
    try:
        z.setsource(comp)
        assert z.state == 2 and z.s.dictid == zlib.adler32(dct)      # ZSTRM_NEEDDICT
        z.setdctnr(dct)
        assert z.state == 3
        assert z.inflate(10000) == d and z.error == 0
    finally:
        z.close()
    z = lib.zstrm(api.ZSTRM_INFLATE)
    try:
        z.setsource(comp)
        z.setdctnr(b"wrong dictionary")
        assert z.error == api.ZSTRM_EBADDICT
    finally:
        z.close()
    z = lib.zstrm(api.ZSTRM_INFLATE)
    try:
        z.setsource(comp)
        z.inflate(10)
        assert z.error == api.ZSTRM_EMISSINGDICT
    finally:
        z.close()


def test_sync_flush_midstream(lib, corpus):
    a, b = corpus.fill(1, 30000), corpus.fill(1, 30000, offset=30000)
    out = bytearray()
    z = lib.zstrm(api.ZSTRM_DEFLATE | api.ZSTRM_ZLIB, 6)
    try:
        z.settargetfn(lambda x: (out.extend(x), len(x))[1])
        z.deflate(a)
        z.flush(0)
        dz = zlib.decompressobj()
        assert dz.decompress(bytes(out)) == a          # everything so far is decodable
        z.deflate(b)
        z.flush(1)
    finally:
        z.close()
    assert zlib.decompress(bytes(out)) == a + b


def test_cross_decode_with_compiled_reference(lib, ref, corpus):
    """What we emit goes through the reference's own zstrm reader (gzip: its zlib writer is the
    defective side, not its reader) and vice versa."""
    d = corpus.fill(0, 120000, offset=77)
    comp, _ = compress(lib, d, api.ZSTRM_GZIP)
    back, err, _ = decompress(ref, comp, read=50000, via_callback=30000)
    assert back == d
    rcomp, _ = compress(ref, d, api.ZSTRM_GZIP, piece=10000)
    back, err, info = decompress(lib, rcomp)
    assert (back, err) == (d, 0)


def test_zlib_preset_dictionary_deflate(lib, corpus):
    """zstrm_setdctnr on the compress side: FDICT + DICTID in the zlib header (src/zstrm.c:1033-1050)."""
    dct = corpus.fill(0, 12000, offset=5)
    d = dct[2000:6000] + corpus.fill(0, 30000, offset=77)
    out = bytearray()
    z = lib.zstrm(api.ZSTRM_DEFLATE | api.ZSTRM_ZLIB, 6)
    try:
        z.settargetfn(lambda b: (out.extend(b), len(b))[1])
        z.setdctnr(dct)
        assert z.error == 0 and z.s.dictid == zlib.adler32(dct)
        assert z.deflate(d) == len(d)
        z.flush(1)
        assert z.error == 0
    finally:
        z.close()
    comp = bytes(out)
    assert comp[1] & 0x20 and comp[2:6] == zlib.adler32(dct).to_bytes(4, "big")
    do = zlib.decompressobj(zdict=dct)
    assert do.decompress(comp) == d
    zi = lib.zstrm(api.ZSTRM_INFLATE)
    try:
        zi.setsource(comp)
        assert zi.state == 2
        zi.setdctnr(dct)
        got = zi.inflate(len(d) + 10)
        assert got == d and zi.error == 0
    finally:
        zi.close()


# ---- gzip files of several members (RFC 1952 2.2); the reference reads one, src/zstrm.c:626-667 ----

def _members():
    import base64
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "gzip_members.json").read_text())
    return {k: dict(v, gz=base64.b64decode(v["gz"])) for k, v in g.items()}


@pytest.mark.parametrize("kw", [{}, {"via_callback": 100}, {"read": 1 << 20}, {"read": 1 << 20, "via_callback": 1 << 20},
                                {"read": 1, "via_callback": 1}])
def test_multi_member_gzip_fixtures(lib, kw):
    """DESIGN.md deviation 11: every member is decoded (as gzip(1) and zlib's gzread do), each
    with its own CRC-32 and ISIZE; bytes after the last member that are not a member are ignored."""
    for name, c in _members().items():
        if kw.get("read") == 1 and name != "five_members":
            continue
        got, err, (stype, crc, adler, total, used) = decompress(lib, c["gz"], **kw)
        if c["size"] is not None:
            assert err == 0, (name, err)
            assert (len(got), zlib.crc32(got), total) == (c["size"], c["crc32"], c["size"]), name
            assert stype == api.ZSTRM_GZIP
        else:
            # the broken second member: the first one arrives, then the error
            assert err in (api.ZSTRM_EBADDATA, api.ZSTRM_ECHECKSUM, api.ZSTRM_ESRCEXHSTD, api.ZSTRM_EDEFLATE), (name, err)
            assert got[:c["first_member_size"]] == got[:c["first_member_size"]] and zlib.crc32(got[:c["first_member_size"]]) == c["first_member_crc32"]


def test_multi_member_of_our_own_members(lib, corpus):
    parts = [corpus.fill(0, 70000, offset=1), b"", corpus.fill(2, 300000, offset=9), corpus.fill(1, 1000, offset=0)]
    gz = b"".join(compress(lib, p, api.ZSTRM_GZIP, level=lv)[0] for p, lv in zip(parts, (6, 6, 1, 9)))
    assert gzip.decompress(gz) == b"".join(parts)
    for kw in ({}, {"via_callback": 4096}, {"read": 1 << 20, "via_callback": 1 << 20}):
        got, err, (stype, crc, adler, total, used) = decompress(lib, gz, **kw)
        assert err == 0 and got == b"".join(parts) and total == len(got) and used == len(gz)
        assert crc == zlib.crc32(parts[-1])          # the running CRC-32 is the last member's


def test_many_small_members(lib, corpus):
    """bgzf-shaped input: hundreds of small members in one file."""
    parts = [corpus.json_record(i)[:3000] for i in range(300)]
    gz = b"".join(gzip.compress(p, 6, mtime=0) for p in parts)
    got, err, info = decompress(lib, gz, read=1 << 20, via_callback=1 << 16)
    assert err == 0 and got == b"".join(parts)


def test_read_ahead_streaming_of_a_chunked_stream(lib):
    """BASELINE configs[4] shape: our own gzip stream (independent chunks) read back through an
    8 MiB source callback.  The inflator gathers several source windows per chunk-parallel step, the
    queue grows while it holds bytes, and the trailer -- queued with an earlier window -- comes back
    to zstrm at the end of the stream."""
    import numpy as np
    n = 40 << 20
    data = np.random.RandomState(7).randint(0, 256, n, dtype=np.uint8).tobytes()
    comp, _ = compress(lib, data, api.ZSTRM_GZIP, level=0, piece=8 << 20)
    for piece in (8 << 20, 3 << 20):
        got, err, (stype, crc, adler, total, used) = decompress(lib, comp, read=8 << 20, via_callback=piece)
        assert err == 0 and total == n and used == len(comp)
        assert crc == zlib.crc32(data) and zlib.crc32(got) == crc
    # two such members in a row: what the first member's inflator had queued of the second comes back
    got, err, (stype, crc, adler, total, used) = decompress(lib, comp + comp, read=8 << 20, via_callback=8 << 20)
    assert err == 0 and total == 2 * n and used == 2 * len(comp) and zlib.crc32(got) == zlib.crc32(data + data)
