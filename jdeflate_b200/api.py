"""ctypes mirror of the jdeflate C API (``include/jdeflate/*.h``).

One :class:`JDeflateLib` wraps one shared object that exports the jdeflate
symbols.  The product library (``jdeflate_b200/lib/libjdeflate.so``) is what
:func:`load` returns; the test-suite also binds the compiled reference
(``oracle/_ref/libjdeflate_ref.so``) through the very same class, which is what
makes the parity tests read like calls into the reference.

The header-inline accessors of the C API (``deflator_setsrc`` & co, reference
jdeflate/deflator.h:159-203, jdeflate/inflator.h:145-189) are reproduced here by
poking the public struct fields, exactly like the C inlines do.

Pointers are plain integers: a host address (``ctypes.addressof``, numpy
``.ctypes.data``) or a device address (``torch.Tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import lib_path

# eDEFLTResult / eINFLTResult
OK, SRCEXHSTD, TGTEXHSTD, ERROR = 0, 1, 2, 3
# eDEFLTFlush
DEFLT_NOFLUSH, DEFLT_END, DEFLT_FLUSH = 0, 1, 2
DEFLT_FIXEDCODES = 0x01
# eDEFLTError
DEFLT_EBADSTATE, DEFLT_EOOM, DEFLT_ELEVEL, DEFLT_EINCORRECTUSE = 1, 2, 3, 4
# eINFLTError
(INFLT_EBADSTATE, INFLT_EBADCODE, INFLT_EBADTREE, INFLT_EFAROFFSET, INFLT_EBADBLOCK,
 INFLT_EINPUTEND, INFLT_EOOM, INFLT_EINCORRECTUSE) = range(1, 9)
# zstrm flags
ZSTRM_INFLATE, ZSTRM_DEFLATE = 0x00010000, 0x00020000
ZSTRM_DFLT, ZSTRM_ZLIB, ZSTRM_GZIP = 0x00100000, 0x00200000, 0x00400000
ZSTRM_DOCRC, ZSTRM_DOADLER, ZSTRM_NOCRC, ZSTRM_NOADLER = 0x01000000, 0x02000000, 0x04000000, 0x08000000
(ZSTRM_OK, ZSTRM_EIOERROR, ZSTRM_EOOM, ZSTRM_EBADDATA, ZSTRM_ECHECKSUM, ZSTRM_EFORMAT,
 ZSTRM_EMISSINGDICT, ZSTRM_ESRCEXHSTD, ZSTRM_ETGTEXHSTD, ZSTRM_EDEFLATE, ZSTRM_EBADDICT,
 ZSTRM_ELIMIT, ZSTRM_EINCORRECTUSE) = range(13)

POISON = 0xDEADBEEF


class TCodec(C.Structure):
    """Public block shared by TDeflator and TInflator (72 bytes on LP64)."""
    _fields_ = [
        ("state", C.c_uint32), ("error", C.c_uint32), ("flags", C.c_uint32),
        ("flush", C.c_uint32),          # `finalinput` in TInflator
        ("status", C.c_uint32),
        ("source", C.c_void_p), ("sbgn", C.c_void_p), ("send", C.c_void_p),
        ("target", C.c_void_p), ("tbgn", C.c_void_p), ("tend", C.c_void_p),
    ]


class TZStrm(C.Structure):
    _fields_ = [
        ("state", C.c_uint32), ("error", C.c_uint32), ("flags", C.c_uint32),
        ("smode", C.c_uint32), ("stype", C.c_uint32), ("level", C.c_int32),
        ("total", C.c_size_t), ("dictid", C.c_uint32), ("dict", C.c_uint32),
        ("crc", C.c_uint32), ("adler", C.c_uint32), ("usedinput", C.c_size_t),
    ]


class JDVersion(C.Structure):
    _fields_ = [("major", C.c_int), ("minor", C.c_int), ("patch", C.c_int),
                ("versionstring", C.c_char_p), ("builddate", C.c_char_p)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint64), ("ms", C.c_double)]


class BatchItem(C.Structure):
    """TJDB200Item: byte ranges relative to the batch base pointers."""
    _fields_ = [("srcoffset", C.c_uint64), ("tgtoffset", C.c_uint64), ("srcsize", C.c_uint64), ("tgtsize", C.c_uint64)]


class BatchResult(C.Structure):
    """TJDB200Result."""
    _fields_ = [("status", C.c_uint32), ("error", C.c_uint32), ("zerror", C.c_uint32), ("checksum", C.c_uint32),
                ("srcused", C.c_uint64), ("tgtused", C.c_uint64)]


JDB200_RAW, JDB200_ZLIB = 0, 1

IFN = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_size_t, C.c_void_p)
OFN = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_size_t, C.c_void_p)

assert C.sizeof(TCodec) == 72 and C.sizeof(TZStrm) == 56


def _addr(buf) -> int:
    """Address of a bytes-like / ctypes / numpy object, or the int itself."""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return C.addressof((C.c_char * len(buf)).from_buffer(buf)) if not isinstance(buf, bytes) \
            else C.cast(C.c_char_p(buf), C.c_void_p).value
    return C.addressof(buf)


class JDeflateLib:
    """All 24 jdeflate entry points of one shared object."""

    SYMBOLS = [
        "deflator_create", "deflator_destroy", "deflator_deflate", "deflator_setdctnr", "deflator_reset",
        "inflator_create", "inflator_destroy", "inflator_inflate", "inflator_setdctnr", "inflator_reset",
        "zstrm_create", "zstrm_destroy", "zstrm_setsource", "zstrm_setsourcefn", "zstrm_settargetfn",
        "zstrm_setdctnr", "zstrm_inflate", "zstrm_deflate", "zstrm_flush", "zstrm_reset",
        "zstrm_crc32update", "zstrm_adler32update", "jdeflate_getversion",
    ]

    def __init__(self, path):
        self.path = str(path)
        if not Path(self.path).exists():
            raise ImportError(
                f"{self.path} is missing: build it first (python -m jdeflate_b200.build); "
                "jdeflate_b200 has no CPU fallback")
        lib = self.lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        P = C.POINTER
        self.missing = []

        def bind(name, restype, argtypes):
            try:
                fn = getattr(lib, name)
            except AttributeError:
                self.missing.append(name)
                return
            fn.restype = restype
            fn.argtypes = argtypes

        bind("deflator_create", P(TCodec), [C.c_size_t, C.c_ssize_t, C.c_void_p])
        bind("deflator_destroy", None, [P(TCodec)])
        bind("deflator_deflate", C.c_int, [P(TCodec), C.c_int])
        bind("deflator_setdctnr", None, [P(TCodec), C.c_void_p, C.c_size_t])
        bind("deflator_reset", None, [P(TCodec)])
        bind("inflator_create", P(TCodec), [C.c_size_t, C.c_void_p])
        bind("inflator_destroy", None, [P(TCodec)])
        bind("inflator_inflate", C.c_int, [P(TCodec), C.c_uint32])
        bind("inflator_setdctnr", None, [P(TCodec), C.c_void_p, C.c_size_t])
        bind("inflator_reset", None, [P(TCodec)])
        bind("zstrm_create", P(TZStrm), [C.c_size_t, C.c_ssize_t, C.c_void_p])
        bind("zstrm_destroy", None, [P(TZStrm)])
        bind("zstrm_setsource", None, [P(TZStrm), C.c_void_p, C.c_size_t])
        bind("zstrm_setsourcefn", None, [P(TZStrm), IFN, C.c_void_p])
        bind("zstrm_settargetfn", None, [P(TZStrm), OFN, C.c_void_p])
        bind("zstrm_setdctnr", None, [P(TZStrm), C.c_void_p, C.c_size_t])
        bind("zstrm_inflate", C.c_size_t, [P(TZStrm), C.c_void_p, C.c_size_t])
        bind("zstrm_deflate", C.c_size_t, [P(TZStrm), C.c_void_p, C.c_size_t])
        bind("zstrm_flush", None, [P(TZStrm), C.c_uint32])
        bind("zstrm_reset", None, [P(TZStrm)])
        bind("zstrm_crc32update", C.c_uint32, [C.c_uint32, C.c_void_p, C.c_size_t])
        bind("zstrm_adler32update", C.c_uint32, [C.c_uint32, C.c_void_p, C.c_size_t])
        bind("jdeflate_getversion", JDVersion, [])
        # the reference object lacks zstrm_crc32combine (SURVEY defect 4) and
        # exports crc32_ncombine instead; the product exports both
        bind("zstrm_crc32combine", C.c_uint32, [C.c_uint32, C.c_uint32, C.c_size_t])
        bind("crc32_ncombine", C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32])

        # additive B200 entry points (include/jdeflate/b200.h); absent from the reference object
        bind("jdb200_inflate_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int])
        bind("jdb200_deflate_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_ssize_t])
        bind("jdb200_set_device", C.c_int, [C.c_int])
        bind("jdb200_device_count", C.c_int, [])
        bind("jdb200_last_error", C.c_char_p, [])
        bind("jdb200_profile", C.c_int, [C.c_int])
        bind("jdb200_profile_read", C.c_int, [C.POINTER(KernelStat), C.c_int])

    def has(self, name: str) -> bool:
        return hasattr(self.lib, name)

    # ---- checksums ------------------------------------------------------
    def crc32(self, data, n=None, value=0) -> int:
        """zlib-style finalised CRC-32 through zstrm_crc32update."""
        n = len(data) if n is None else n
        if n == 0:
            return value
        return self.lib.zstrm_crc32update(value ^ 0xFFFFFFFF, _addr(data), n) ^ 0xFFFFFFFF

    def adler32(self, data, n=None, value=1) -> int:
        n = len(data) if n is None else n
        if n == 0:
            return value
        return self.lib.zstrm_adler32update(value, _addr(data), n)

    def crc32_combine(self, c1, c2, len2) -> int:
        if self.has("zstrm_crc32combine"):
            return self.lib.zstrm_crc32combine(c1, c2, len2)
        return self.lib.crc32_ncombine(c1, c2, len2)

    def version(self) -> str:
        return self.lib.jdeflate_getversion().versionstring.decode()

    # ---- additive B200 calls -------------------------------------------------
    def profile(self, enable=True):
        rc = self.lib.jdb200_profile(1 if enable else 0)
        if rc:
            raise RuntimeError(f"jdb200_profile rc={rc}")

    def profile_read(self) -> dict:
        """{kernel name: (launches, summed ms)}"""
        arr = (KernelStat * 32)()
        n = self.lib.jdb200_profile_read(arr, 32)
        return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].ms)) for i in range(n)}

    def inflate_batch(self, source, target, items, results, count, fmt=JDB200_RAW) -> int:
        """All four arguments are addresses (host or device); see include/jdeflate/b200.h."""
        return self.lib.jdb200_inflate_batch(_addr(source), _addr(target), _addr(items), _addr(results), count, fmt)

    def inflate_batch_bytes(self, streams, sizes, fmt=JDB200_RAW):
        """Host convenience: decode a list of byte strings; returns (outputs, results)."""
        n = len(streams)
        items = (BatchItem * n)()
        so = to = 0
        for i, (z, cap) in enumerate(zip(streams, sizes)):
            items[i] = BatchItem(so, to, len(z), cap)
            so += len(z)
            to += cap
        src = C.create_string_buffer(b"".join(streams), max(so, 1))
        dst = C.create_string_buffer(max(to, 1))
        res = (BatchResult * n)()
        rc = self.lib.jdb200_inflate_batch(C.addressof(src), C.addressof(dst), C.addressof(items), C.addressof(res), n, fmt)
        if rc:
            raise RuntimeError(f"jdb200_inflate_batch rc={rc}: {self.lib.jdb200_last_error().decode()}")
        outs = [dst.raw[items[i].tgtoffset: items[i].tgtoffset + res[i].tgtused] for i in range(n)]
        return outs, list(res)

    def deflate_batch(self, source, target, items, results, count, fmt=JDB200_RAW, level=6) -> int:
        """All four arguments are addresses (host or device); see include/jdeflate/b200.h."""
        return self.lib.jdb200_deflate_batch(_addr(source), _addr(target), _addr(items), _addr(results), count, fmt, level)

    def deflate_batch_bytes(self, records, caps=None, fmt=JDB200_RAW, level=6):
        """Host convenience: compress a list of byte strings, each into its own stream;
        returns (streams, results).  ``caps``: target bytes per record (default: always enough)."""
        n = len(records)
        items = (BatchItem * n)()
        so = to = 0
        for i, r in enumerate(records):
            cap = caps[i] if caps is not None else len(r) + len(r) // 64 + 80
            items[i] = BatchItem(so, to, len(r), cap)
            so += len(r)
            to += cap
        src = C.create_string_buffer(b"".join(records), max(so, 1))
        dst = C.create_string_buffer(max(to, 1))
        res = (BatchResult * n)()
        rc = self.lib.jdb200_deflate_batch(C.addressof(src), C.addressof(dst), C.addressof(items), C.addressof(res), n, fmt, level)
        if rc:
            raise RuntimeError(f"jdb200_deflate_batch rc={rc}: {self.lib.jdb200_last_error().decode()}")
        outs = [dst.raw[items[i].tgtoffset: items[i].tgtoffset + res[i].tgtused] for i in range(n)]
        return outs, list(res)

    # ---- object factories ----------------------------------------------
    def deflator(self, level=6, flags=0):
        return Deflator(self, level, flags)

    def inflator(self, flags=0):
        return Inflator(self, flags)

    def zstrm(self, flags, level=0):
        return ZStrm(self, flags, level)

    # ---- one-shot conveniences on host memory ---------------------------
    def deflate_bytes(self, data: bytes, level=6, flags=0, flush=DEFLT_END, cap=None,
                      window=None, feed=None) -> bytes:
        """Compress ``data`` like the README loop of the reference.

        ``window``: target window size per call (None = one big window);
        ``feed``: source piece size per setsrc (None = all at once).
        """
        d = self.deflator(level, flags)
        try:
            return d.run(data, flush=flush, cap=cap, window=window, feed=feed)
        finally:
            d.close()

    def inflate_bytes(self, data: bytes, cap: int, window=None, feed=None, final=True):
        """Returns (status, error, output bytes, consumed)."""
        s = self.inflator()
        try:
            return s.run(data, cap, window=window, feed=feed, final=final)
        finally:
            s.close()


class _Codec:
    def __init__(self, jd, ptr):
        if not ptr:
            raise MemoryError("create() returned NULL (bad argument, no CUDA device or out of memory)")
        self.jd = jd
        self.p = ptr
        self.s = ptr.contents
        self._keep = []

    # the header-inline accessors ---------------------------------------
    def setsrc(self, addr, size):
        assert size > 0
        if self.s.flush:
            if self.s.error == 0:
                self.s.error = self._EINCORRECTUSE
                self.s.state = POISON
            return
        a = _addr(addr)
        self._keep.append(addr)
        self.s.source = self.s.sbgn = a
        self.s.send = a + size

    def settgt(self, addr, size):
        assert size > 0
        a = _addr(addr)
        self._keep.append(addr)
        self.s.target = self.s.tbgn = a
        self.s.tend = a + size

    def srcend(self) -> int:
        return (self.s.source or 0) - (self.s.sbgn or 0)

    def tgtend(self) -> int:
        return (self.s.target or 0) - (self.s.tbgn or 0)

    @property
    def error(self):
        return self.s.error

    @property
    def status(self):
        return self.s.status

    @property
    def state(self):
        return self.s.state


class Deflator(_Codec):
    _EINCORRECTUSE = DEFLT_EINCORRECTUSE

    def __init__(self, jd, level=6, flags=0):
        super().__init__(jd, jd.lib.deflator_create(flags, level, None))
        self.level = level

    def deflate(self, flush=DEFLT_NOFLUSH) -> int:
        return self.jd.lib.deflator_deflate(self.p, flush)

    def setdctnr(self, d: bytes):
        self.jd.lib.deflator_setdctnr(self.p, _addr(d), len(d))

    def reset(self):
        self.jd.lib.deflator_reset(self.p)

    def close(self):
        if self.p:
            self.jd.lib.deflator_destroy(self.p)
            self.p = None

    def run(self, data, flush=DEFLT_END, cap=None, window=None, feed=None) -> bytes:
        n = len(data)
        cap = cap if cap is not None else n + n // 8 + 4096
        out = bytearray()
        win = C.create_string_buffer(window if window else cap)
        src = C.create_string_buffer(bytes(data), max(n, 1))
        feed = feed or max(n, 1)
        pos = 0
        while True:
            piece = min(feed, n - pos)
            last = pos + piece >= n
            if piece:
                self.setsrc(C.addressof(src) + pos, piece)
            pos += piece
            f = flush if last else DEFLT_NOFLUSH
            if piece == 0 and not last:
                raise RuntimeError("no progress")
            if piece == 0 and n == 0:
                # nothing to feed at all: the reference still wants a source
                self.setsrc(C.addressof(src), 1)
                self.s.send = self.s.sbgn
            while True:
                self.settgt(win, len(win))
                r = self.deflate(f)
                out += win.raw[: self.tgtend()]
                if r != TGTEXHSTD:
                    break
            if r == ERROR:
                raise RuntimeError(f"deflator error {self.error}")
            if r == OK or (last and r == SRCEXHSTD and f == DEFLT_NOFLUSH):
                return bytes(out)
            if last and r != SRCEXHSTD:
                return bytes(out)


class Inflator(_Codec):
    _EINCORRECTUSE = INFLT_EINCORRECTUSE

    def __init__(self, jd, flags=0):
        super().__init__(jd, jd.lib.inflator_create(flags, None))

    def inflate(self, final=0) -> int:
        return self.jd.lib.inflator_inflate(self.p, final)

    def setdctnr(self, d: bytes):
        self.jd.lib.inflator_setdctnr(self.p, _addr(d), len(d))

    def reset(self):
        self.jd.lib.inflator_reset(self.p)

    def close(self):
        if self.p:
            self.jd.lib.inflator_destroy(self.p)
            self.p = None

    def run(self, data, cap, window=None, feed=None, final=True):
        """Decode with the README double loop; returns (status, error, out, consumed)."""
        n = len(data)
        src = C.create_string_buffer(bytes(data), max(n, 1))
        out = bytearray()
        win = C.create_string_buffer(window if window else max(cap, 1))
        feed = feed or max(n, 1)
        pos = 0
        consumed = 0
        r = SRCEXHSTD
        while True:
            piece = min(feed, n - pos)
            if piece == 0:
                break
            last = pos + piece >= n
            self.setsrc(C.addressof(src) + pos, piece)
            while True:
                self.settgt(win, len(win))
                r = self.inflate(1 if (last and final) else 0)
                out += win.raw[: self.tgtend()]
                if r != TGTEXHSTD or len(out) >= cap + (window or 0):
                    break
            consumed = pos + self.srcend()
            pos += piece
            if r != SRCEXHSTD:
                break
        return r, self.error, bytes(out), consumed


class ZStrm:
    def __init__(self, jd, flags, level=0):
        self.jd = jd
        self.p = jd.lib.zstrm_create(flags, level, None)
        if not self.p:
            raise MemoryError("zstrm_create returned NULL")
        self.s = self.p.contents
        self._keep = []

    def close(self):
        if self.p:
            self.jd.lib.zstrm_destroy(self.p)
            self.p = None

    def settargetfn(self, pyfn):
        """pyfn(bytes) -> int bytes written (or negative)."""
        def tramp(buf, size, user):
            return pyfn(C.string_at(buf, size))
        cb = OFN(tramp)
        self._keep.append(cb)
        self.jd.lib.zstrm_settargetfn(self.p, cb, None)

    def setsourcefn(self, pyfn):
        """pyfn(maxsize) -> bytes (b'' at EOF) or negative int."""
        def tramp(buf, size, user):
            r = pyfn(size)
            if isinstance(r, int):
                return r
            C.memmove(buf, r, len(r))
            return len(r)
        cb = IFN(tramp)
        self._keep.append(cb)
        self.jd.lib.zstrm_setsourcefn(self.p, cb, None)

    def setsource(self, data: bytes):
        self._keep.append(data)
        self.jd.lib.zstrm_setsource(self.p, _addr(data), len(data))

    def setdctnr(self, d: bytes):
        self._keep.append(d)
        self.jd.lib.zstrm_setdctnr(self.p, _addr(d), len(d))

    def deflate(self, data, n=None) -> int:
        n = len(data) if n is None else n
        self._keep.append(data)
        return self.jd.lib.zstrm_deflate(self.p, _addr(data), n)

    def inflate(self, n: int) -> bytes:
        buf = C.create_string_buffer(max(n, 1))
        got = self.jd.lib.zstrm_inflate(self.p, buf, n)
        return buf.raw[:got]

    def inflate_into(self, addr, n: int) -> int:
        return self.jd.lib.zstrm_inflate(self.p, _addr(addr), n)

    def flush(self, final=1):
        self.jd.lib.zstrm_flush(self.p, final)

    def reset(self):
        self.jd.lib.zstrm_reset(self.p)

    @property
    def error(self):
        return self.s.error

    @property
    def state(self):
        return self.s.state


_default = None


def load(path=None) -> JDeflateLib:
    """The product library.  Raises ImportError when it has not been built."""
    global _default
    if path is not None:
        return JDeflateLib(path)
    if _default is None:
        _default = JDeflateLib(lib_path())
    return _default
