"""Threading contract of the boundary: distinct instances are fully concurrent, there is
no hidden global state a second thread could trip over (the reference has none either,
src/deflator.c:266-268).  Four host threads, each owning a TDeflator, a TInflator and a TZStrm,
run at the same time on the B200; every result must be bit-exact and equal to what the same work
gives single-threaded.  ctypes releases the GIL for the duration of a foreign call, so the library
calls of the threads really overlap."""
import threading
import zlib

import pytest

from jdeflate_b200 import api

pytestmark = pytest.mark.gpu
NTHREADS = 4
ROUNDS = 3


def _work(jd, data: bytes, level: int):
    """deflator -> inflator -> zstrm gzip -> zstrm gunzip, returns everything it produced."""
    comp = jd.deflate_bytes(data, level)
    st, err, back, used = jd.inflate_bytes(comp, len(data))
    assert (st, err, used) == (api.OK, 0, len(comp))
    assert back == data
    out = bytearray()
    zs = jd.zstrm(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, level)
    zs.settargetfn(lambda b: (out.extend(b), len(b))[1])
    half = len(data) // 2
    assert zs.deflate(data[:half]) == half
    assert zs.deflate(data[half:]) == len(data) - half
    zs.flush(1)
    assert zs.error == 0
    zs.close()
    gz = bytes(out)
    zi = jd.zstrm(api.ZSTRM_INFLATE | api.ZSTRM_GZIP)
    zi.setsource(gz)
    again = zi.inflate(len(data) + 64)
    assert zi.error == 0
    zi.close()
    assert again == data
    crc = jd.crc32(data)
    return comp, gz, crc


def test_four_threads_own_instances(jd, corpus):
    inputs = [corpus.fill(k % 3 if k % 3 != 1 else 2, (3 << 20) + 12345 * k, offset=k << 20) for k in range(NTHREADS)]
    levels = [6, 1, 9, 6]
    single = [_work(jd, inputs[k], levels[k]) for k in range(NTHREADS)]
    for k in range(NTHREADS):
        assert zlib.decompress(single[k][0], -15) == inputs[k]
        assert zlib.decompress(single[k][1], 31) == inputs[k]
        assert single[k][2] == zlib.crc32(inputs[k])

    results = [[None] * ROUNDS for _ in range(NTHREADS)]
    errors = []
    barrier = threading.Barrier(NTHREADS)

    def run(k):
        try:
            for r in range(ROUNDS):
                barrier.wait()
                results[k][r] = _work(jd, inputs[k], levels[k])
        except BaseException as e:      # noqa: BLE001 -- report whatever a thread dies of
            errors.append((k, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=run, args=(k,)) for k in range(NTHREADS)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(NTHREADS):
        for r in range(ROUNDS):
            assert results[k][r] == single[k], (k, r)


def test_threads_share_the_stateless_helpers(jd, corpus):
    """zstrm_crc32update / zstrm_adler32update and the batch entry points keep per-thread scratch."""
    data = [corpus.fill(0, (2 << 20) + k, offset=k << 18) for k in range(NTHREADS)]
    want = [(zlib.crc32(d), zlib.adler32(d)) for d in data]
    comp = [[zlib.compress(d[i << 16:(i + 1) << 16], 6) for i in range(16)] for d in data]
    got = [None] * NTHREADS
    errors = []

    def run(k):
        try:
            for _ in range(4):
                c, a = jd.crc32(data[k]), jd.adler32(data[k])
                outs, res = jd.inflate_batch_bytes(comp[k], [1 << 16] * 16, fmt=api.JDB200_ZLIB)
                assert b"".join(outs) == data[k][: 16 << 16]
                got[k] = (c, a)
        except BaseException as e:      # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=run, args=(k,)) for k in range(NTHREADS)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert got == want
