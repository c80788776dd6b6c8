"""Host wall-clock per API call of one resident Q1 step (bench.py's chain op), WITHOUT extra synchronisation: shows where
the step's non-kernel time goes (operator creation, the synchronising calls, result export, teardown)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sail_b200 import engine

ctx = engine.Context(0)
table = bench.gen_shard(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0, 0, 1).combine_chunks()
fused, final, sort = bench.q1_specs()
dev = engine.to_device(table, ctx)
chain = {"op": "chain", "ops": [fused, final, sort]}
T = {}
N = 30
for it in range(N + 3):
    if it == 3:
        T.clear(); ctx.synchronize(); t_all = time.perf_counter()
    marks = [("start", time.perf_counter())]
    op = engine.GpuExec(chain, [table.schema], ctx); marks.append(("create", time.perf_counter()))
    b = dev.borrow(); marks.append(("borrow", time.perf_counter()))
    op.push(b); marks.append(("push", time.perf_counter()))
    op.finish(); marks.append(("finish", time.perf_counter()))
    out = op.collect(); marks.append(("collect", time.perf_counter()))
    mm = op.metrics(); marks.append(("metrics", time.perf_counter()))
    op.close(); marks.append(("close", time.perf_counter()))
    for (_, a), (n, b_) in zip(marks, marks[1:]):
        T[n] = T.get(n, 0.0) + (b_ - a) * 1e3
ctx.synchronize()
total = (time.perf_counter() - t_all) * 1e3 / N
for k, v in T.items():
    print(f"{k:10s} {v / N:8.3f} ms")
print(f"{'sum':10s} {sum(T.values()) / N:8.3f} ms   wall/step {total:8.3f} ms   pipeline kernels/step {mm['gpu.pipeline_kernel_ns'] / 1e6:.3f} ms")
del dev
