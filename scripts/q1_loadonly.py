"""How fast can the specialised tile pipeline STREAM Q1's seven columns (100 B/row) when there is next to nothing to compute?
Filter on the date and the two flag columns (never false) + a keyless sum of the four decimals: same TMA traffic as Q1, ~1/4 of
its instructions.  If this kernel reaches the copy roofline, Q1's 0.61 is compute; if it stops near 4 TB/s, it is the load path
(4 KB bulk copies from seven streams).  usage: q1_loadonly.py [orders]   (env: SAILGPU_RPT, SAILGPU_JIT_STAGES)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SAILGPU_TIMING", "1")
os.environ.setdefault("SAILGPU_JIT_MIN_ROWS", "0")
import bench  # noqa: E402
from datagen import tpch_gpu  # noqa: E402
from sail_b200 import engine, plans  # noqa: E402

n_orders = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 15_000_000
C = {n: i for i, n in enumerate(bench.Q1_COLS)}
col = lambda n: {"col": C[n]}
pred = plans.and_(plans.binop("<=", col("l_shipdate"), plans.date("1998-12-01")),
                  plans.binop("!=", col("l_returnflag"), plans.string("Z")), plans.binop("!=", col("l_linestatus"), plans.string("Z")))
spec = {"op": "pipeline", "stages": [
    {"op": "filter", "predicate": pred, "projection": None},
    {"op": "aggregate", "mode": "partial", "group_by": [],
     "aggs": [{"fn": "sum", "args": [col(c)], "name": c, "input_type": "Decimal128(15,2)"} for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")]
             + [{"fn": "count", "args": [], "name": "c", "input_type": None}]}]}
if "--source" in sys.argv:
    import pyarrow as pa
    D = pa.decimal128(15, 2)
    sch = pa.schema([("l_quantity", D), ("l_extendedprice", D), ("l_discount", D), ("l_tax", D), ("l_returnflag", pa.string_view()), ("l_linestatus", pa.string_view()), ("l_shipdate", pa.date32())])
    n, src = engine.jit_precompile(spec, [sch], 0, 0)
    print(src[:200], "...", n)
    sys.exit(0)
ctx = engine.default_context()
gen = tpch_gpu.generate_buffers(10.0, 0, n_orders, (), bench.Q1_COLS, 0)[1]
dev = gen.device_batch(ctx)
out = {"rows": gen.rows, "rpt": os.environ.get("SAILGPU_RPT"), "stages": os.environ.get("SAILGPU_JIT_STAGES")}
for rep in range(4):
    op = engine.GpuExec(spec, [gen.schema], ctx)
    op.push(dev.borrow())
    op.finish()
    res = op.collect()
    m = op.metrics()
    op.close()
    out[f"kernel_ms_{rep}"] = round(m["gpu.pipeline_kernel_ns"] / 1e6, 3)
    out["jit_launches"] = m.get("gpu.jit_launches")
ms = out["kernel_ms_3"]
out["GBps"] = round(gen.rows * 100 / ms / 1e6, 1) if ms else None
print(json.dumps(out), flush=True)
