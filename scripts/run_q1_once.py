"""Runs the resident Q1 plan a few times over one GPU-generated batch (for ncu captures).  usage: run_q1_once.py [sf<=10] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from datagen import tpch, tpch_gpu  # noqa: E402
from sail_b200 import engine  # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = engine.Context(0)
gen = tpch_gpu.generate_buffers(sf, 0, tpch.counts(sf)["orders"], (), bench.Q1_COLS, 0)[1]
dev = gen.device_batch(ctx)
specs = bench.q1_specs()
for _ in range(reps):
    out, *_ = bench.run_query(ctx, specs, [dev], gen.schema)
print(out.to_pylist()[0])
del dev
ctx.synchronize()
