"""oracle/cpipelines.py -- TEST INFRASTRUCTURE / CPU BASELINE.  ctypes front end of oracle/cpipelines.c
(the C restatement of the reference's CPU execution of Q1 / Q6).  Consumes pyarrow columns in place."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_pipelines.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpipelines.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-pthread", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.q1_cpu.restype = ctypes.c_int
        _lib.q6_cpu.restype = ctypes.c_int64
    return _lib


def _addr(col, buf=1):
    """address of the values buffer of a single-chunk column at its offset"""
    arr = col.chunk(0) if isinstance(col, pa.ChunkedArray) else col
    assert not isinstance(col, pa.ChunkedArray) or col.num_chunks == 1
    b = arr.buffers()[buf]
    width = {pa.date32(): 4}.get(arr.type, 16)
    return ctypes.c_void_p(b.address + arr.offset * width)


def _i128(a: np.ndarray, i: int) -> int:
    lo, hi = int(a[2 * i]), int(a[2 * i + 1])
    return (np.int64(hi).item() << 64) | lo


def q1(lineitem: pa.Table, cutoff_days: int, threads: int = 1):
    """-> list of (returnflag, linestatus, sum_qty, sum_base_price, sum_disc_price, sum_charge, sum_disc, count)
    with decimals as unscaled ints (scales 2,2,4,6,2)."""
    t = lineitem.combine_chunks()
    n = t.num_rows
    keys = np.zeros(64 * 32, dtype=np.uint8)
    sums = np.zeros(64 * 5 * 2, dtype=np.uint64)
    counts = np.zeros(64, dtype=np.int64)
    g = lib().q1_cpu(ctypes.c_int64(n), _addr(t["l_quantity"]), _addr(t["l_extendedprice"]), _addr(t["l_discount"]),
                     _addr(t["l_tax"]), _addr(t["l_returnflag"]), _addr(t["l_linestatus"]), _addr(t["l_shipdate"]),
                     ctypes.c_int32(cutoff_days), ctypes.c_int(threads), keys.ctypes.data_as(ctypes.c_void_p),
                     sums.ctypes.data_as(ctypes.c_void_p), counts.ctypes.data_as(ctypes.c_void_p))
    out = []
    for i in range(g):
        k = keys[32 * i: 32 * i + 32].tobytes()
        rf = k[4:4 + int.from_bytes(k[0:4], "little")].decode()
        ls = k[20:20 + int.from_bytes(k[16:20], "little")].decode()
        out.append((rf, ls) + tuple(_i128(sums, 5 * i + j) for j in range(5)) + (int(counts[i]),))
    return sorted(out)


def q6(lineitem: pa.Table, d0: int, d1: int, threads: int = 1):
    t = lineitem.combine_chunks()
    s = np.zeros(2, dtype=np.uint64)
    rows = lib().q6_cpu(ctypes.c_int64(t.num_rows), _addr(t["l_quantity"]), _addr(t["l_extendedprice"]), _addr(t["l_discount"]),
                        _addr(t["l_shipdate"]), ctypes.c_int32(d0), ctypes.c_int32(d1), ctypes.c_int(threads),
                        s.ctypes.data_as(ctypes.c_void_p))
    return rows, _i128(s, 0)
