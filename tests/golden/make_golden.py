"""Generate tests/golden/golden.json from the COMPILED REFERENCE (oracle/_ref/libjdeflate_ref.so,
built by oracle/Makefile from the unmodified sources under /root/reference) and zlib 1.3.

Run in the build container only (the GPU box has no /root/reference and only consumes the
committed JSON):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures are what pins
the CPU oracle (oracle/jd_oracle.c) and, through it, the CUDA path:

  inflate_kat   hand-assembled raw DEFLATE streams (valid and malformed) with the status / error /
                output / consumed-bytes the reference's inflator_inflate reports for them
  deflate_ref   size + CRC-32 of the reference deflator's output for corpus slices at levels 0-9
                (the oracle's jdo_deflate must reproduce these streams byte for byte)
  checksums     zlib crc32 / adler32 of corpus slices, and the reference's zstrm_crc32update
  corpus        CRC-32 of the synthetic corpus generators (pins tools/corpus.c)
"""
import base64
import json
import sys
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from jdeflate_b200 import api  # noqa: E402
from support import BitWriter, Corpus, KIND_NAMES, zlib_raw  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "libjdeflate_ref.so"


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def kat_streams():
    """(name, stream bytes, output capacity, final flag)"""
    out = []
    W = BitWriter

    # --- valid streams -------------------------------------------------------
    out.append(("empty_stored_final", W().put(1, 1).put(0, 2).align().raw(b"\x00\x00\xff\xff").bytes(), 16, 1))
    out.append(("empty_fixed_final", W().put(1, 1).put(1, 2).fixed_lit(256).bytes(), 16, 1))
    out.append(("single_literal_fixed", W().put(1, 1).put(1, 2).fixed_lit(ord("a")).fixed_lit(256).bytes(), 16, 1))
    out.append(("stored_hello", W().put(1, 1).put(0, 2).align().raw(b"\x05\x00\xfa\xffhello").bytes(), 16, 1))
    # two stored blocks, the first not final, bit padding in between
    out.append(("stored_two_blocks",
                W().put(0, 1).put(0, 2).align().raw(b"\x03\x00\xfc\xffabc")
                .put(1, 1).put(0, 2).align().raw(b"\x02\x00\xfd\xffde").bytes(), 16, 1))
    # fixed block: 'a' then a match of length 258 at distance 1 (code 285, dist code 0)
    w = W().put(1, 1).put(1, 2).fixed_lit(ord("a")).fixed_lit(285).huff(0, 5).fixed_lit(256)
    out.append(("fixed_len258_dist1", w.bytes(), 300, 1))
    # fixed block: 4 literals, match length 3 distance 4, length 10 (code 264) distance 7 (code 5, 1 extra bit)
    w = W().put(1, 1).put(1, 2)
    for ch in b"abcd":
        w.fixed_lit(ch)
    w.fixed_lit(257).huff(3, 5)                     # len 3, dist 4
    w.fixed_lit(264).huff(5, 5).put(0, 1)           # len 10, dist 7
    w.fixed_lit(256)
    out.append(("fixed_matches", w.bytes(), 64, 1))
    # distance 32768 (code 29, 13 extra bits all ones) after 32768 stored bytes
    big = bytes((i * 7 + (i >> 8)) & 0xff for i in range(32768))
    w = W().put(0, 1).put(0, 2).align().raw(b"\x00\x80\xff\x7f" + big)
    w.put(1, 1).put(1, 2).fixed_lit(258).huff(29, 5).put(0x1fff, 13).fixed_lit(256)
    out.append(("dist_32768", w.bytes(), 40000, 1))
    # zlib-made dynamic blocks (third party encoder)
    c = Corpus()
    for kind, n in ((0, 3000), (2, 5000), (4, 9000), (3, 700)):
        data = c.fill(kind, n, offset=12345)
        for lvl in (1, 6, 9):
            out.append((f"zlib_{KIND_NAMES[kind]}_{n}_L{lvl}", zlib_raw(data, lvl), n + 16, 1))
    # trailing garbage after the final block
    out.append(("trailing_garbage", zlib_raw(c.fill(0, 2000), 6) + b"GARBAGE!", 2100, 1))
    # 15-bit codes: a skewed distribution forces long codes
    skew = bytearray()
    for i in range(40):
        skew += bytes([i]) * max(1, 2 ** max(0, 14 - i) // 8)
    out.append(("long_codes", zlib_raw(bytes(skew), 9), len(skew) + 16, 1))
    # a dynamic block with a single distance code (incomplete distance tree of one 1-bit code)
    out.append(("single_dist_code", zlib_raw(b"ab" + b"x" * 600 + b"cd", 6), 700, 1))

    # --- malformed -------------------------------------------------------------
    out.append(("btype3", W().put(1, 1).put(3, 2).put(0, 13).bytes(), 16, 1))
    out.append(("stored_len_mismatch", W().put(1, 1).put(0, 2).align().raw(b"\x05\x00\xfa\xfehello").bytes(), 16, 1))
    # match reaching before the start of the output
    w = W().put(1, 1).put(1, 2).fixed_lit(ord("a")).fixed_lit(257).huff(4, 5).put(0, 1).fixed_lit(256)
    out.append(("far_offset", w.bytes(), 64, 1))
    good = zlib_raw(c.fill(0, 4000, offset=777), 6)
    out.append(("truncated_final", good[: len(good) // 2], 4100, 1))
    out.append(("truncated_not_final", good[: len(good) // 2], 4100, 0))
    out.append(("target_too_small", good, 1000, 1))
    # dynamic header with HLIT = 31 (288 > 286 codes)
    out.append(("bad_hlit", W().put(1, 1).put(2, 2).put(31, 5).put(0, 5).put(0, 4).put(0, 12).put(0, 32).bytes(), 16, 1))
    # dynamic header whose first length code is a repeat (16) with nothing before it
    w = W().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(15, 4)
    # precode lengths in order 16,17,18,0,8,...: give 16 and 0 one bit each
    pl = {16: 1, 0: 1}
    for s in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15):
        w.put(pl.get(s, 0), 3)
    w.huff(1, 1).put(0, 2)            # canonical: 0 -> code 0, 16 -> code 1 ; emit symbol 16
    w.put(0, 32)
    out.append(("repeat_without_previous", w.bytes(), 16, 1))
    # over-subscribed precode: three symbols of length 1
    w = W().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(15, 4)
    pl = {16: 1, 17: 1, 18: 1}
    for s in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15):
        w.put(pl.get(s, 0), 3)
    w.put(0, 32)
    out.append(("oversubscribed_precode", w.bytes(), 16, 1))
    # fixed block using the reserved length symbol 286
    out.append(("reserved_litlen_286", W().put(1, 1).put(1, 2).fixed_lit(286).huff(0, 5).fixed_lit(256).bytes(), 64, 1))
    # fixed block using the reserved distance code 30
    out.append(("reserved_dist_30", W().put(1, 1).put(1, 2).fixed_lit(ord("a")).fixed_lit(257).huff(30, 5).fixed_lit(256).bytes(), 64, 1))
    out.append(("empty_input_final", b"", 16, 1))
    return out


def main():
    if not REF.exists():
        raise SystemExit(f"{REF} missing: run `make -C oracle ref` in the build container first")
    ref = api.JDeflateLib(REF)
    c = Corpus()
    g = {"reference_version": ref.version(), "zlib_version": zlib.ZLIB_VERSION}

    kats = []
    for name, stream, cap, final in kat_streams():
        if len(stream) == 0:
            # the reference asserts on an empty source window: the expectation is by contract
            kats.append({"name": name, "stream": "", "cap": cap, "final": final, "status": api.ERROR,
                         "error": api.INFLT_EINPUTEND, "out": "", "ref_consumed": 0, "exact_consumed": 0,
                         "by_contract": True})
            continue
        st, err, out, used = ref.inflate_bytes(stream, cap, final=bool(final))
        exact = None
        try:
            d = zlib.decompressobj(-15)
            zout = d.decompress(stream)
            if d.eof:
                exact = len(stream) - len(d.unused_data)
                assert zout[: len(out)] == out or st != api.OK
        except zlib.error:
            pass
        kats.append({"name": name, "stream": b64(stream), "cap": cap, "final": final, "status": st, "error": err,
                     "out_len": len(out), "out_crc": zlib.crc32(out), "out": b64(out) if len(out) <= 512 else None,
                     "ref_consumed": used, "exact_consumed": exact})
    g["inflate_kat"] = kats

    defl = []
    for kind in range(5):
        for n, off in ((0, 0), (1, 0), (2, 5), (3, 9), (4, 11), (257, 0), (4096, 77), (65536, 1000), (300000, 31)):
            data = c.fill(kind, n, offset=off) if n else b""
            for lvl in (0, 1, 2, 5, 6, 7, 9):
                z = ref.deflate_bytes(data, lvl) if n else None
                if z is None:
                    continue
                assert zlib.decompress(z, -15) == data
                defl.append({"kind": kind, "n": n, "offset": off, "level": lvl, "size": len(z), "crc": zlib.crc32(z)})
    # fixed-codes flag
    for kind in (0, 2):
        data = c.fill(kind, 50000, offset=3)
        z = ref.deflate_bytes(data, 6, flags=api.DEFLT_FIXEDCODES)
        defl.append({"kind": kind, "n": 50000, "offset": 3, "level": 6, "flags": 1, "size": len(z), "crc": zlib.crc32(z)})
    g["deflate_ref"] = defl

    cks = []
    for kind in range(5):
        for n, off in ((1, 0), (7, 3), (4096, 0), (5551, 9), (5552, 9), (5553, 9), (65521, 1), (1000003, 17)):
            data = c.fill(kind, n, offset=off)
            cks.append({"kind": kind, "n": n, "offset": off, "crc": zlib.crc32(data), "adler": zlib.adler32(data),
                        "ref_crc": ref.crc32(data)})
            assert cks[-1]["crc"] == cks[-1]["ref_crc"]
    g["checksums"] = cks
    g["crc_combine"] = [{"c1": zlib.crc32(b"hello "), "c2": zlib.crc32(b"world"), "len2": 5, "crc": zlib.crc32(b"hello world")},
                        {"c1": zlib.crc32(c.fill(0, 70000)), "c2": zlib.crc32(c.fill(2, 123457)), "len2": 123457,
                         "crc": zlib.crc32(c.fill(0, 70000) + c.fill(2, 123457))}]

    g["corpus"] = [{"kind": k, "n": 1 << 20, "offset": off, "crc": zlib.crc32(c.fill(k, 1 << 20, offset=off))}
                   for k in range(6) for off in (0, (4 << 20) - 1000)]
    g["json_records"] = [{"index": i, "size": len(c.json_record(i)), "crc": zlib.crc32(c.json_record(i))} for i in (0, 1, 2, 1000, 65535)]

    path = Path(__file__).resolve().parent / "golden.json"
    path.write_text(json.dumps(g, indent=1))
    print(path, path.stat().st_size, "bytes;", len(kats), "inflate KATs,", len(defl), "deflate refs")


if __name__ == "__main__":
    main()
