"""Runs the resident Q1 plan a few times (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sail_b200 import engine  # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = engine.Context(0)
table = bench.gen_shard(sf, 0, 1).combine_chunks()
specs = bench.q1_specs()
bench.SORT_ON_GPU = False
dev = engine.to_device(table, ctx)
for _ in range(reps):
    out, *_ = bench.run_query(ctx, specs, [dev], table.schema)
print(out.to_pylist()[0])
del dev
ctx.synchronize()
