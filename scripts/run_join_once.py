"""Runs the PK-FK hash join orders(o_orderkey, o_orderdate, o_shippriority) x lineitem(l_orderkey, l_extendedprice, l_discount) a few
times over GPU-generated tables (15 M x 60 M rows at SF10) -- for ncu captures and timing.  usage: run_join_once.py [sf<=10] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datagen import tpch, tpch_gpu  # noqa: E402
from sail_b200 import engine  # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = engine.Context(0)
o, l = tpch_gpu.generate_buffers(sf, 0, tpch.counts(sf)["orders"], ["o_orderkey", "o_orderdate", "o_shippriority"], ["l_orderkey", "l_extendedprice", "l_discount"], 0)
do, dl = o.device_batch(ctx), l.device_batch(ctx)
spec = {"op": "hash_join", "join_type": "inner", "on": [[0, 0]], "filter": None, "projection": [1, 2, 3, 4, 5]}
for r in range(reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    op = engine.GpuExec(spec, [o.schema, l.schema], ctx)
    op.push(do.borrow(), 0)
    op.finish(0)
    ctx.synchronize()
    t1 = time.perf_counter()
    op.push(dl.borrow(), 1)
    op.finish(1)
    out = op.collect_device()
    ctx.synchronize()
    t2 = time.perf_counter()
    m = op.metrics()
    rows = sum(d.num_rows for d in out)
    op.close()
    print(f"rep {r}: build {(t1 - t0) * 1e3:.2f} ms, probe {(t2 - t1) * 1e3:.2f} ms, out rows {rows}, launches {m['gpu.kernel_launches']}", flush=True)
    del out
del do, dl
ctx.synchronize()
