#!/usr/bin/env python
"""bench.py -- TPC-H Q1 (scan + filter + hash-aggregate) at SF10 per GPU: BASELINE.json configs[1].

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference ...                      the reference's CPU algorithm on the host cores
  torchrun ... bench.py --gpus N ...                        one rank per GPU, lineitem sharded by order range

A "step" is one execution of the Q1 physical plan (fused FilterExec+ProjectionExec+AggregateExec
partial -> AggregateExec final -> SortExec) over the rank's lineitem shard.
  value : rows/s with the Arrow column buffers already resident in HBM (device batches pushed zero-copy)
  e2e   : rows/s through the same C ABI with HOST (pinned) Arrow buffers: H2D copies and the D2H of
          the result are inside the timed region
  roofline : the dominant kernel (pipeline_kernel) -- algorithmic bytes (100 B/row)
             / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : oracle/cpipelines.c (C port of the reference's CPU path) on all host cores
Inputs (6 GB per GPU) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("SAILGPU_TIMING", "1")

Q1_COLS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
ALGO_BYTES_PER_ROW = 100      # 4 x Decimal128 + 2 x Utf8View + Date32  
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sf", type=float, default=10.0, help="scale factor PER GPU (weak scaling)")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 22, help="rows per host batch in the e2e leg")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gen_shard(sf_per_gpu: float, rank: int, world: int):
    """lineitem rows of orders [rank*O, (rank+1)*O) of an SF(sf*world) database"""
    from datagen import tpch
    total_sf = sf_per_gpu * world
    orders_total = tpch.counts(total_sf)["orders"]
    per = orders_total // world
    return tpch.lineitem(total_sf, Q1_COLS, first=rank * per, n=per)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, every 2 ms; nvidia-smi as fallback)."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.sm, self.bits, self.stop_flag, self.thread, self.max_mhz = index, [], 0, False, None, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.sm.append(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    self.bits |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    time.sleep(0.002)
                else:
                    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    self.bits |= int(out[2].strip(), 16)
            except Exception:
                time.sleep(0.01)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": [n for n, b in self.REASONS.items() if self.bits & b],
                "samples": len(sm)}


def q1_specs():
    from sail_b200 import plans
    final_sorted = plans.q1()
    final = final_sorted.inputs[0]
    partial = final.inputs[0]
    stages, n = [], partial
    while n.spec["op"] in ("filter", "projection", "aggregate"):
        stages.append(n.spec)
        n = n.inputs[0]
    return {"op": "pipeline", "stages": stages[::-1]}, final.spec, final_sorted.spec


def pin_table(ctx, table):
    """copy every buffer of a single-chunk table into pinned host memory; returns (pinned table, keepalive, bytes)"""
    import pyarrow as pa
    from sail_b200 import engine
    keep, arrays, total = [], [], 0
    for col in table.columns:
        arr = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
        bufs = []
        for b in arr.buffers():
            if b is None:
                bufs.append(None)
                continue
            p = ctypes.c_void_p()
            rc = engine.lib().sailgpu_host_alloc(ctx._h, b.size, ctypes.byref(p))
            assert rc == 0
            ctypes.memmove(p.value, b.address, b.size)
            keep.append(p)
            bufs.append(pa.foreign_buffer(p.value, b.size, base=p))
            total += b.size
        arrays.append(pa.Array.from_buffers(arr.type, len(arr), bufs, null_count=arr.null_count, offset=arr.offset))
    return pa.table(arrays, names=table.schema.names), keep, total


def run_query_dist(backend, specs, inputs, in_schema, host_chunks=None):
    """world > 1: fused partial on the shard -> hash repartition + NCCL all-to-all -> final on the owner ->
    gather to rank 0 -> sort there.  Returns (result table on rank 0, launches, kernel ns, kernel launches)."""
    from sail_b200 import dist as sdist
    from sail_b200 import engine
    fused, final, sort = specs
    if not os.environ.get("SAILGPU_DIST_PYTHON"):
        # the whole plan, exchanges included, as one chain inside the library (the step-by-step driver below is kept for
        # comparison: SAILGPU_DIST_PYTHON=1)
        op = engine.GpuExec(sdist.two_phase_chain(fused, final, [0, 1], [sort]), [in_schema], backend.ctx)
        if host_chunks is None:
            for d in inputs:
                op.push(d.borrow())
        else:
            for b in host_chunks:
                op.push(b)
        op.finish()
        out = op.collect()
        mm = op.metrics()
        op.close()
        return (out if backend.rank == 0 else None), mm["gpu.kernel_launches"], mm["gpu.pipeline_kernel_ns"], mm["gpu.pipeline_launches"]
    op1 = engine.GpuExec(fused, [in_schema], backend.ctx)
    if host_chunks is None:
        for d in inputs:
            op1.push(d.borrow())
    else:
        for b in host_chunks:
            op1.push(b)
    op1.finish()
    parts = op1.collect_device()
    m1 = op1.metrics()
    pschema = op1.schema
    op1.close()
    for p in parts:
        p.schema = pschema
    backend.launches = 0
    fin, on_root = sdist.final_aggregate(backend, parts, pschema, [0, 1], final)
    if not on_root:
        fin = sdist.gather_to_root(backend, fin, fin[0].schema)
    table = backend.run_to_host(sort, fin) if backend.rank == 0 else None
    return table, m1["gpu.kernel_launches"] + backend.launches, m1["gpu.pipeline_kernel_ns"], m1["gpu.pipeline_launches"]


def run_query(ctx, specs, inputs, in_schema, host_chunks=None):
    """One Q1 execution.  inputs: list of DeviceBatch (resident leg) or None with host_chunks (e2e leg).
    Returns (result table, #kernel launches, pipeline kernel ns, pipeline launches)."""
    from sail_b200 import engine
    if DIST_BACKEND is not None:
        return run_query_dist(DIST_BACKEND, specs, inputs, in_schema, host_chunks)
    fused, final, sort = specs
    # one GPU island: fused partial -> final -> sort handed over inside the library (HBM), result pulled to the host
    chain = {"op": "chain", "ops": [fused, final] + ([sort] if SORT_ON_GPU else [])}
    op = engine.GpuExec(chain, [in_schema], ctx)
    if host_chunks is None:
        for d in inputs:
            op.push(d.borrow())
    else:
        for b in host_chunks:
            op.push(b)
    op.finish()
    out = op.collect()
    mm = op.metrics()
    op.close()
    return out, mm["gpu.kernel_launches"], mm["gpu.pipeline_kernel_ns"], mm["gpu.pipeline_launches"]


SORT_ON_GPU = True
DIST_BACKEND = None


def check_result(table, want_rows):
    """GPU result vs the C oracle (decimals compared as unscaled integers)."""
    import decimal
    got = []
    for r in table.to_pylist():
        def u(v, s):
            return int(decimal.Decimal(v).scaleb(s))
        got.append((r["l_returnflag"], r["l_linestatus"], u(r["sum_qty"], 2), u(r["sum_base_price"], 2), u(r["sum_disc_price"], 4),
                    u(r["sum_charge"], 6), r["count_order"]))
    want = [(w[0], w[1], w[2], w[3], w[4], w[5], w[7]) for w in want_rows]
    assert sorted(got) == sorted(want), f"GPU Q1 result differs from the CPU oracle:\n{sorted(got)}\n{sorted(want)}"


def cpu_q1(table, threads, reps):
    from oracle import cpipelines
    from sail_b200 import plans
    cutoff = plans.days("1998-09-24")
    cpipelines.q1(table.slice(0, min(table.num_rows, 1 << 20)), cutoff, threads)      # warm up / page in
    t0 = time.perf_counter()
    for _ in range(reps):
        rows = cpipelines.q1(table, cutoff, threads)
    dt = (time.perf_counter() - t0) / reps
    return rows, dt


def e2e_leg(args, ctx, specs, table, stream, world, total_rows):
    import torch
    import torch.distributed as dist
    n_rows = table.num_rows
    pinned, keep, h2d_bytes = pin_table(ctx, table)
    chunks = [pinned.slice(o, args.e2e_chunk).to_batches()[0] for o in range(0, n_rows, args.e2e_chunk)]
    for _ in range(max(1, min(2, args.warmup))):
        out_e, _, _, _ = run_query(ctx, specs, None, table.schema, host_chunks=chunks)
    if world > 1:
        dist.barrier()
    ctx.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 5))
    e2.record(stream)
    for _ in range(e2e_steps):
        out_e, _, _, _ = run_query(ctx, specs, None, table.schema, host_chunks=chunks)
    e3.record(stream)
    ctx.synchronize()
    ms_e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms_e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = float(t.item())
    e2e_value = total_rows / (ms_e / e2e_steps / 1e3)
    d2h_bytes = 0 if out_e is None else sum(b.size for c in out_e.columns for ch in c.chunks for b in ch.buffers() if b is not None)
    return e2e_value, ms_e, e2e_steps, h2d_bytes, d2h_bytes, chunks, out_e


def main():
    args = parse_args()
    rank, world, local = dist_env()
    # libraries (NCCL, torchrun) may print to stdout: keep fd 1 for the ONE JSON line only
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if world > 1 and args.gpus != world:
        args.gpus = world
    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args)

    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sail_b200 import engine
    global SORT_ON_GPU, DIST_BACKEND
    ctx = engine.Context(local)
    if world > 1:
        from sail_b200 import dist as sdist
        uid = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        DIST_BACKEND = sdist.GpuBackend(ctx, rank, world)
    table = gen_shard(args.sf, rank, world).combine_chunks()
    n_rows = table.num_rows
    specs = q1_specs()
    try:
        engine.GpuExec(specs[2], [engine.GpuExec(specs[1], [engine.GpuExec(specs[0], [table.schema], ctx).schema], ctx).schema], ctx).close()
    except engine.SailGpuError:
        SORT_ON_GPU = False
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    # ---- resident leg ---------------------------------------------------------------------------
    dev = engine.to_device(table, ctx)
    for _ in range(max(3, args.warmup)):
        out, _, _, _ = run_query(ctx, specs, [dev], table.schema)
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    ctx.synchronize()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = kern_ns = kern_launches = 0
    for _ in range(args.steps):
        out, l, kns, kl = run_query(ctx, specs, [dev], table.schema)
        launches += l
        kern_ns += kns
        kern_launches += kl
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    total_rows = n_rows * world
    value = total_rows / (ms_per_step / 1e3)

    # ---- end-to-end leg: host (pinned) Arrow buffers through the C ABI -------------------------------
    e2e_value = ms_e = None
    e2e_steps = 1
    h2d_bytes = d2h_bytes = 0
    chunks = []
    out_e = out
    if not args.skip_e2e:
        e2e_value, ms_e, e2e_steps, h2d_bytes, d2h_bytes, chunks, out_e = e2e_leg(args, ctx, specs, table, stream, world, total_rows)
    # ---- CPU baseline (rank 0, N=1 only) + parity of the full-size GPU result ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        want_rows, dt = cpu_q1(table, threads, reps=3)
        check_result(out, want_rows)
        if not args.skip_e2e:
            check_result(out_e, want_rows)
        cpu = {"value": n_rows / dt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"full SF{args.sf:g} lineitem ({n_rows} rows) x3, oracle/cpipelines.c (C port of the DataFusion CPU path), all host threads"}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = FALLBACK_HBM_GBS, "fallback"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        # the fused stage may run as several launches per step (cardinality-probe chunk + the rest): bytes of all its
        # launches over the time of all its launches, i.e. the byte-weighted average launch
        kern_ms = kern_ns / 1e6 / max(1, args.steps)
        achieved = (n_rows * ALGO_BYTES_PER_ROW) / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "q1_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": "TPC-H Q1 rows/s (scan+filter+hash-aggregate), lineitem resident in HBM",
            "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i128 (Decimal128) / i64", "data": "synthetic (dbgen-exact TPC-H lineitem)",
            "config": {"workload": f"TPC-H Q1 SF{args.sf:g} per GPU, 1 partition per GPU, Arrow batches resident in HBM",
                       "rows_per_gpu": n_rows, "strings": "Utf8View", "l2": "inputs (6 GB) larger than L2; no flush",
                       "plan": "GpuChainExec{GpuPipelineExec[Filter+Projection+Aggregate(Partial)] -> GpuAggregateExec(FinalPartitioned)"
                               + (" -> GpuExchangeExec(auto: coalesce on rank 0 | Hash + NCCL all-to-all)" if world > 1 else "")
                               + (" -> GpuSortExec" if SORT_ON_GPU else "") + "}", "parallelism": f"{world} rank(s), lineitem sharded by order range"},
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                                                    "ms_per_step": ms_e / e2e_steps, "host_batches": len(chunks)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "kernel": "sg::pipeline_kernel", "kernel_ms": kern_ms, "peak_source": peak_src,
                         "launches_per_step": kern_launches / max(1, args.steps), "algorithmic_bytes_per_launch": n_rows * ALGO_BYTES_PER_ROW},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    del dev, out, out_e, chunks
    ctx.synchronize()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm for this path (C port: the Rust
    toolchain and DataFusion are absent from this image), all host threads, same config/metric."""
    table = gen_shard(args.sf, 0, 1).combine_chunks()
    n_rows = table.num_rows
    threads = os.cpu_count() or 1
    from oracle import cpipelines
    from sail_b200 import plans
    cutoff = plans.days("1998-09-24")
    for _ in range(max(1, args.warmup)):
        cpipelines.q1(table, cutoff, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpipelines.q1(table, cutoff, threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_rows / dt
    line = {"impl": "reference", "metric": "TPC-H Q1 rows/s (scan+filter+hash-aggregate), lineitem resident in HBM",
            "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i128 (Decimal128)",
            "data": "synthetic (dbgen-exact TPC-H lineitem)",
            "config": {"workload": f"TPC-H Q1 SF{args.sf:g}, Arrow batches resident in host memory", "rows": n_rows},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": f"full SF{args.sf:g} lineitem ({n_rows} rows) per step"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
