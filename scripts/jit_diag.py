"""Specialised-kernel diagnostics: the fused Q1 pipeline over GPU-generated lineitem at growing sizes, each size in its own
process under a timeout (a kernel that does not terminate must not take the whole GPU call with it).
usage: python scripts/jit_diag.py [orders ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(n_orders: int):
    import bench
    from datagen import tpch_gpu
    from sail_b200 import engine
    ctx = engine.default_context()
    fused, final, sort = bench.q1_specs()
    gen = tpch_gpu.generate_buffers(10.0, 0, n_orders, (), bench.Q1_COLS, 0)[1]
    dev = gen.device_batch(ctx)
    out = {"orders": n_orders, "rows": gen.rows}
    for rep in range(3):
        t0 = time.perf_counter()
        op = engine.GpuExec({"op": "chain", "ops": [fused, final, sort]}, [gen.schema], ctx)
        op.push(dev.borrow())
        op.finish()
        res = op.collect()
        m = op.metrics()
        op.close()
        out[f"ms_{rep}"] = round((time.perf_counter() - t0) * 1e3, 3)
        out[f"kernel_ms_{rep}"] = round(m["gpu.pipeline_kernel_ns"] / 1e6, 3)
        out["jit_launches"] = m.get("gpu.jit_launches")
    out["groups"] = res.num_rows
    out["count"] = sum(res.column("count_order").to_pylist())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(int(sys.argv[2]))
        sys.exit(0)
    sizes = [int(a) for a in sys.argv[1:]] or [20_000, 250_000, 1_000_000, 4_000_000, 15_000_000]
    for n in sizes:
        for env in ({}, {"SAILGPU_JIT": "0"}):
            e = dict(os.environ, SAILGPU_TIMING="1", SAILGPU_JIT_MIN_ROWS="0", **env)
            try:
                r = subprocess.run([sys.executable, __file__, "--one", str(n)], env=e, capture_output=True, text=True, timeout=150)
                print(("jit " if not env else "vm  ") + (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"{'jit' if not env else 'vm '} orders={n}: TIMEOUT (150 s)", flush=True)
