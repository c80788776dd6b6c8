"""Runs GROUP BY l_orderkey (sum(l_quantity), count(*)) a few times over one GPU-generated lineitem batch: 60 M rows -> 15 M groups
at SF10, the high-cardinality case of the aggregate (for ncu captures and timing).  usage: run_agg_once.py [sf<=10] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datagen import tpch, tpch_gpu  # noqa: E402
from sail_b200 import engine  # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = sys.argv[3] if len(sys.argv) > 3 else "dec"      # dec: sum(Decimal128) + count; cnt: count only; i64: sum(Int64) + count
ctx = engine.Context(0)
gen = tpch_gpu.generate_buffers(sf, 0, tpch.counts(sf)["orders"], (), ["l_orderkey", "l_quantity"], 0)[1]
dev = gen.device_batch(ctx)
aggs = {"dec": [{"fn": "sum", "args": [{"col": 1}], "name": "sum_qty"}, {"fn": "count", "args": [], "name": "cnt"}],
        "cnt": [{"fn": "count", "args": [], "name": "cnt"}],
        "i64": [{"fn": "sum", "args": [{"col": 0}], "name": "sum_key"}, {"fn": "count", "args": [], "name": "cnt"}]}[variant]
spec = {"op": "aggregate", "mode": "single", "group_by": [{"expr": {"col": 0}, "name": "l_orderkey"}], "aggs": aggs}
for r in range(reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    op = engine.GpuExec(spec, [gen.schema], ctx)
    op.push(dev.borrow())
    op.finish()
    out = op.collect_device()
    ctx.synchronize()
    dt = time.perf_counter() - t0
    m = op.metrics()
    rows = sum(d.num_rows for d in out)
    op.close()
    print(f"[{variant}] rep {r}: {dt * 1e3:.2f} ms, groups {rows}, launches {m['gpu.kernel_launches']} (specialised {m['gpu.jit_launches']}), pipeline kernels {m['gpu.pipeline_kernel_ns'] / 1e6:.2f} ms", flush=True)
    del out
del dev
ctx.synchronize()
