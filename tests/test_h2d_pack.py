"""CPU tests of the host loops of the packed ingest (sail_b200/csrc/h2d_pack.cpp): the one-pass "pack under an assumed encoding and
report what the piece really holds" functions against numpy, over the AVX-512 / AVX2 / baseline clones glibc picks on this machine.
The GPU side (expansion kernels, the pool) is covered by every GPU test that pushes a host batch."""
import ctypes

import numpy as np
import pytest

from sail_b200 import engine


@pytest.fixture(scope="module")
def L():
    lib = engine.lib()
    lib.sg_packchk_dec128.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.sg_packchk_dec128.restype = None
    lib.sg_packchk_views.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32]
    lib.sg_packchk_views.restype = ctypes.c_uint32
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 1000, 65536 + 5])
@pytest.mark.parametrize("w,dt", [(1, np.uint8), (2, np.uint16), (4, np.uint32)])
def test_packchk_dec128_stores_deltas_and_reports_the_true_range(L, n, w, dt):
    rng = np.random.default_rng(n * 7 + w)
    lo = rng.integers(-3000, 5001, n).astype(np.int64)
    a = np.zeros(2 * n + 2, dtype=np.int64)
    a[0:2 * n:2] = lo
    a[1:2 * n:2] = lo >> 63                      # high word = sign extension (what a Decimal128(15,2) looks like)
    out = np.zeros(n * 4 + 64, dtype=np.uint8)
    mn, mx, bad = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_uint64()
    base = -4000
    L.sg_packchk_dec128(ptr(out), ptr(a), n, base, w, ctypes.byref(mn), ctypes.byref(mx), ctypes.byref(bad))
    got = out[: n * w].view(dt).astype(np.int64)
    assert (got == ((lo - base) & ((1 << (8 * w)) - 1))).all()
    assert bad.value == 0
    if n:
        assert (mn.value, mx.value) == (int(lo.min()), int(lo.max()))
        a[2 * (n // 2) + 1] = 7                  # one value that does not fit in 64 bits: must be reported
        L.sg_packchk_dec128(ptr(out), ptr(a), n, base, w, ctypes.byref(mn), ctypes.byref(mx), ctypes.byref(bad))
        assert bad.value != 0


@pytest.mark.parametrize("n", [0, 1, 15, 16, 33, 4096 + 3])
@pytest.mark.parametrize("maxlen", [0, 1, 5, 12])
def test_packchk_views_packs_inline_views_and_reports_the_longest(L, n, maxlen):
    rng = np.random.default_rng(n + 31 * maxlen)
    lens = rng.integers(0, maxlen + 1, n).astype(np.uint32)
    if n:
        lens[n // 2] = maxlen
    v = np.zeros((n + 1, 16), dtype=np.uint8)
    body = rng.integers(65, 91, (n, 12)).astype(np.uint8)
    for i in range(n):
        v[i, 0:4] = np.frombuffer(np.uint32(lens[i]).tobytes(), dtype=np.uint8)
        v[i, 4:4 + lens[i]] = body[i, : lens[i]]
    out = np.zeros((n + 4) * (1 + maxlen) + 64, dtype=np.uint8)
    got_max = L.sg_packchk_views(ptr(out), ptr(v), n, maxlen)
    assert got_max == (int(lens.max()) if n else 0)
    rows = out[: n * (1 + maxlen)].reshape(n, 1 + maxlen) if n else out[:0].reshape(0, 1 + maxlen)
    assert (rows[:, 0] == lens.astype(np.uint8)).all()
    for i in range(n):
        assert (rows[i, 1:1 + lens[i]] == body[i, : lens[i]]).all()
    if n > 2 and maxlen < 12:
        v[1, 0:4] = np.frombuffer(np.uint32(maxlen + 1).tobytes(), dtype=np.uint8)      # a longer value than assumed: the caller re-packs
        assert L.sg_packchk_views(ptr(out), ptr(v), n, maxlen) == maxlen + 1
