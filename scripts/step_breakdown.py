"""Wall-clock breakdown of one resident Q1 step (host-side phases), to find non-kernel overhead."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sail_b200 import engine

ctx = engine.Context(0)
table = bench.gen_shard(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0, 0, 1).combine_chunks()
fused, final, sort = bench.q1_specs()
dev = engine.to_device(table, ctx)
T = {}
def tick(name, t0):
    ctx.synchronize()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()
N = 20
for it in range(N + 3):
    if it == 3:
        T.clear()
    t = time.perf_counter()
    op1 = engine.GpuExec(fused, [table.schema], ctx); t = tick("op1.create", t)
    op1.push(dev.borrow()); t = tick("op1.push(kernel)", t)
    op1.finish(); parts = op1.collect_device(); t = tick("op1.collect_device(extract)", t)
    op2 = engine.GpuExec(final, [op1.schema], ctx); t = tick("op2.create", t)
    for p in parts: op2.push(p)
    t = tick("op2.push", t)
    op2.finish(); mids = op2.collect_device(); t = tick("op2.collect_device", t)
    op3 = engine.GpuExec(sort, [op2.schema], ctx); t = tick("op3.create", t)
    for p in mids: op3.push(p)
    op3.finish(); t = tick("op3.push", t)
    out = op3.collect(); t = tick("op3.collect(sort+D2H)", t)
    for o in (op1, op2, op3): o.metrics(); o.close()
    t = tick("close", t)
tot = 0
for k, v in T.items():
    print(f"{k:32s} {v / N:8.3f} ms"); tot += v / N
print(f"{'total':32s} {tot:8.3f} ms")
del dev
