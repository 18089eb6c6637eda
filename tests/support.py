"""Test-side bindings: the CPU oracle (oracle/_ref/libjd_oracle.so), the synthetic
corpus generator (oracle/_ref/libjd_corpus.so) and small helpers.

Test infrastructure only -- nothing here is imported by the jdeflate_b200 package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "oracle" / "_ref"

TEXT, LOGS, BINARY, INCOMP, JSON, MIXED = range(6)
KIND_NAMES = ["text", "logs", "binary", "incomp", "json", "mixed"]


def _ensure(name: str) -> Path:
    p = OUT / name
    if not p.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port", "tools"], check=True, capture_output=True)
    return p


class Oracle:
    """ctypes view of oracle/jd_oracle.c."""

    def __init__(self):
        lib = self.lib = C.CDLL(str(_ensure("libjd_oracle.so")))
        lib.jdo_crc32_update.restype = C.c_uint32
        lib.jdo_crc32_update.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        lib.jdo_adler32_update.restype = C.c_uint32
        lib.jdo_adler32_update.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        lib.jdo_crc32_combine.restype = C.c_uint32
        lib.jdo_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        lib.jdo_inflate.restype = C.c_int
        lib.jdo_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        lib.jdo_deflate.restype = C.c_int
        lib.jdo_deflate.argtypes = [C.c_int, C.c_uint, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]

    def crc32(self, data: bytes, value=0) -> int:
        return self.lib.jdo_crc32_update(value ^ 0xFFFFFFFF, data, len(data)) ^ 0xFFFFFFFF

    def adler32(self, data: bytes, value=1) -> int:
        return self.lib.jdo_adler32_update(value, data, len(data))

    def crc32_combine(self, c1, c2, n2) -> int:
        return self.lib.jdo_crc32_combine(c1, c2, n2)

    def inflate(self, data: bytes, cap: int, final=True):
        """-> (status, error, out, consumed)"""
        out = C.create_string_buffer(max(cap, 1))
        used, made, err = C.c_size_t(), C.c_size_t(), C.c_int()
        st = self.lib.jdo_inflate(data, len(data), out, cap, 1 if final else 0,
                                  C.byref(used), C.byref(made), C.byref(err))
        return st, err.value, out.raw[: made.value], used.value

    def deflate(self, data: bytes, level=6, flags=0) -> bytes:
        cap = len(data) + len(data) // 8 + 4096
        out = C.create_string_buffer(cap)
        n = C.c_size_t()
        rc = self.lib.jdo_deflate(level, flags, data, len(data), out, cap, C.byref(n))
        if rc != 0:
            raise RuntimeError(f"jdo_deflate rc={rc}")
        return out.raw[: n.value]


class Corpus:
    """ctypes view of tools/corpus.c (deterministic synthetic data)."""

    def __init__(self):
        lib = self.lib = C.CDLL(str(_ensure("libjd_corpus.so")))
        lib.jdc_fill.restype = C.c_int
        lib.jdc_fill.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.jdc_json_record.restype = C.c_int
        lib.jdc_json_record.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32]
        lib.jdc_json_record_size.restype = C.c_uint32
        lib.jdc_json_record_size.argtypes = [C.c_uint64]

    def fill(self, kind: int, n: int, offset: int = 0) -> bytes:
        buf = C.create_string_buffer(max(n, 1))
        assert self.lib.jdc_fill(kind, offset, buf, n) == 0
        return buf.raw[:n]

    def fill_into(self, kind: int, addr: int, n: int, offset: int = 0):
        assert self.lib.jdc_fill(kind, offset, addr, n) == 0

    def json_record(self, index: int, size: int | None = None) -> bytes:
        size = size if size is not None else self.lib.jdc_json_record_size(index)
        buf = C.create_string_buffer(size)
        assert self.lib.jdc_json_record(index, buf, size) == 0
        return buf.raw[:size]


def zlib_raw(data: bytes, level=6) -> bytes:
    z = zlib.compressobj(level, zlib.DEFLATED, -15)
    return z.compress(data) + z.flush()


def zlib_inflate_raw(data: bytes) -> bytes:
    return zlib.decompress(data, -15)


class BitWriter:
    """Hand-assemble DEFLATE bit strings for known-answer / malformed streams."""

    def __init__(self):
        self.bits = []

    def put(self, value, n):            # LSB first (header fields, extra bits)
        for i in range(n):
            self.bits.append((value >> i) & 1)
        return self

    def huff(self, code, n):            # MSB first (Huffman codes)
        for i in range(n - 1, -1, -1):
            self.bits.append((code >> i) & 1)
        return self

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)
        return self

    def raw(self, data: bytes):
        assert len(self.bits) % 8 == 0
        for b in data:
            self.put(b, 8)
        return self

    def bytes(self) -> bytes:
        b = self.bits + [0] * ((-len(self.bits)) % 8)
        return bytes(sum(b[i + j] << j for j in range(8)) for i in range(0, len(b), 8))

    # fixed-code helpers (RFC 1951 3.2.6)
    def fixed_lit(self, sym):
        if sym < 144:
            return self.huff(0x30 + sym, 8)
        if sym < 256:
            return self.huff(0x190 + sym - 144, 9)
        if sym < 280:
            return self.huff(sym - 256, 7)
        return self.huff(0xC0 + sym - 280, 8)
