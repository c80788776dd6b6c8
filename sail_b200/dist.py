"""Multi-GPU driver for the hash-repartitioned part of a plan: one process per GPU.

Mirrors what Sail's cluster mode does with `AggregateExec(Partial) -> RepartitionExec Hash(keys) ->
AggregateExec(FinalPartitioned)` (plan shape: python/pysail/tests/spark/__snapshots__/
test_tpch.plan.yaml:12-14; stage cut at the exchange: crates/sail-execution/src/job_graph/
planner.rs:172-294; channel -> consumer mapping: driver/job_scheduler/core.rs:753-766): every rank
aggregates its shard, hash-partitions the partial states on the group keys into `world` segments,
the segments are exchanged all-to-all (segment p -> rank p) and every rank finalises the groups it
owns.  `gather_to_root` is the `InputMode::Merge` step (final sort / result collection on one rank).

The logic is backend-agnostic so that the CPU test-suite can run it under `gloo` with world_size 2
(`tests/test_dist_cpu.py`): `GpuBackend` executes through libsailgpu + NCCL, `HostBackend` through any
`run_op(spec, *tables)` callable and a torch.distributed object all-to-all.
"""
from __future__ import annotations

import pyarrow as pa


def shard_range(total: int, rank: int, world: int):
    """contiguous row-range shard (like DataSourceExec file groups): returns (first, count)"""
    per, rem = divmod(total, world)
    first = rank * per + min(rank, rem)
    return first, per + (1 if rank < rem else 0)


def repartition_spec(key_cols: list, n: int) -> dict:
    return {"op": "repartition", "scheme": "hash", "exprs": [{"col": c} for c in key_cols], "n": n}


class HostBackend:
    """Reference backend for tests: operators via `run_op`, exchange via torch.distributed (gloo)."""

    def __init__(self, run_op, rank: int, world: int, group=None):
        self.run_op, self.rank, self.world, self.group = run_op, rank, world, group

    def run(self, spec, *tables):
        return self.run_op(spec, *tables)

    def partition(self, spec, table):
        return self.run_op(spec, table)          # list of `world` tables

    def exchange(self, parts: list, schema: pa.Schema):
        import torch.distributed as dist
        out = [None] * self.world
        if self.world == 1:
            out = parts
        else:
            gathered = [None] * self.world
            # gloo has no object all-to-all: all_gather the per-destination lists and pick our column
            dist.all_gather_object(gathered, [p.to_pylist() for p in parts], group=self.group)
            out = [pa.Table.from_pylist(gathered[src][self.rank], schema=schema) for src in range(self.world)]
        return pa.concat_tables(out) if out else schema.empty_table()

    def empty(self, schema: pa.Schema):
        return schema.empty_table()

    def to_host(self, table):
        return table

    def rows(self, table):
        return table.num_rows

    def max_over_ranks(self, value: int) -> int:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def run_to_host(self, spec, *tables):
        return self.run_op(spec, *tables)


class GpuBackend:
    """libsailgpu + NCCL: data stays in HBM between the operators and across the exchange."""

    def __init__(self, ctx, rank: int, world: int):
        from . import engine
        self.engine, self.ctx, self.rank, self.world = engine, ctx, rank, world
        self.launches = 0

    def _drain(self, op):
        parts = op.collect_device()
        self.launches += op.metrics()["gpu.kernel_launches"]
        return parts

    def run(self, spec, *inputs):
        """inputs: lists of DeviceBatch (consumed) or pyarrow tables; returns list of DeviceBatch"""
        e = self.engine
        schemas = [i.schema if isinstance(i, pa.Table) else i[0].schema for i in inputs]
        op = e.GpuExec(spec, schemas, self.ctx)
        for k, i in enumerate(inputs):
            for b in ([i] if isinstance(i, pa.Table) else i):
                op.push(b, k)
            op.finish(k)
        out = self._drain(op)
        schema = op.schema
        op.close()
        for d in out:
            d.schema = schema
        return out

    def partition(self, spec, batches):
        e = self.engine
        op = e.GpuExec(spec, [batches[0].schema], self.ctx)
        for b in batches:
            op.push(b)
        op.finish()
        parts = []
        for p in range(self.world):
            seg = []
            while True:
                d, more = op.pull_device(partition=p)
                if d.num_rows:
                    seg.append(d)
                if not more:
                    break
            parts.append(seg)
        self.launches += op.metrics()["gpu.kernel_launches"]
        schema = op.schema
        op.close()
        # one batch per destination (concatenate through an identity projection when a segment has several)
        return [self._single(seg, schema) for seg in parts]

    def _single(self, seg, schema):
        if not seg:
            return None
        if len(seg) == 1:
            seg[0].schema = schema
            return seg[0]
        ident = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(schema.names)]}
        if not seg:
            return None
        # several batches for one destination: concatenate them on the device (SortExec collects its whole input;
        # row order inside an exchange segment is free)
        op = self.engine.GpuExec({"op": "sort", "keys": [{"expr": {"col": 0}}], "fetch": None}, [schema], self.ctx)
        for b in seg:
            op.push(b)
        op.finish()
        out = op.collect_device()
        op.close()
        out[0].schema = schema
        return out[0]

    def empty(self, schema: pa.Schema):
        return None                      # sailgpu_exchange treats a released / absent batch as empty

    def exchange(self, parts: list, schema: pa.Schema):
        return [self.engine.exchange(parts, schema, self.ctx)]

    def rows(self, batches):
        return sum(b.num_rows for b in batches)

    def max_over_ranks(self, value: int) -> int:
        """agreement on a plan choice: one 8-byte all-reduce on the process group bench.py / the tests initialised"""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return int(value)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def run_to_host(self, spec, *inputs):
        """like run(), but the operator's output is pulled straight into host Arrow memory"""
        e = self.engine
        schemas = [i.schema if isinstance(i, pa.Table) else i[0].schema for i in inputs]
        op = e.GpuExec(spec, schemas, self.ctx)
        for k, i in enumerate(inputs):
            for b in ([i] if isinstance(i, pa.Table) else i):
                op.push(b, k)
            op.finish(k)
        t = op.collect()
        self.launches += op.metrics()["gpu.kernel_launches"]
        op.close()
        return t

    def to_host(self, batches):
        schema = batches[0].schema
        ident = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(schema.names)]}
        op = self.engine.GpuExec(ident, [schema], self.ctx)
        for b in batches:
            op.push(b)
        op.finish()
        t = op.collect()
        op.close()
        return t


def exchange_by_key(backend, data, schema: pa.Schema, key_cols: list):
    """RepartitionExec Hash(key_cols, world) + all-to-all: returns this rank's share"""
    parts = backend.partition(repartition_spec(key_cols, backend.world), data)
    return backend.exchange(parts, schema)


def gather_to_root(backend, data, schema: pa.Schema):
    """InputMode::Merge: everything to rank 0 (other ranks receive nothing)"""
    if backend.world == 1:
        return data
    single = data if isinstance(data, pa.Table) else (data[0] if len(data) == 1 else backend._single(data, schema))
    parts = [single if p == 0 else backend.empty(schema) for p in range(backend.world)]
    return backend.exchange(parts, schema)


# partial-aggregate states of at most this many rows per rank are coalesced on the root instead of hash-partitioned
SMALL_EXCHANGE_ROWS = 1 << 14


def final_aggregate(backend, partial, schema: pa.Schema, key_cols: list, final_spec: dict, small_rows: int = SMALL_EXCHANGE_ROWS):
    """Second phase of a two-phase aggregation across ranks.  Returns (data, on_root): the finalised groups and whether
    they already sit on rank 0 only.

    * many groups: `RepartitionExec Hash(keys)` + all-to-all, `AggregateExec(FinalPartitioned)` on the owner of each
      group (the plan Sail's cluster mode runs, module docstring);
    * a handful of groups on every rank (TPC-H Q1: 4): hash partitioning would cost two exchanges and `world` partition
      pulls for nothing, so the partial states are coalesced on rank 0 (`CoalescePartitionsExec` +
      `AggregateExec(Final)`, what DataFusion plans for a single target partition) -- one exchange.
    The choice must be the same on every rank: it is taken on the maximum partial row count (one 8-byte all-reduce)."""
    if backend.world == 1:
        return backend.run(final_spec, partial), True
    if backend.max_over_ranks(backend.rows(partial)) <= small_rows:
        gathered = gather_to_root(backend, partial, schema)
        if backend.rank != 0:
            return None, True
        return backend.run(final_spec, gathered), True
    mine = exchange_by_key(backend, partial, schema, key_cols)
    return backend.run(final_spec, mine), False


def two_phase_chain(partial_spec: dict, final_spec: dict, key_cols: list, tail: list = (), small_rows: int = SMALL_EXCHANGE_ROWS) -> dict:
    """The whole two-phase aggregation across ranks as ONE operator spec for libsailgpu (`{"op": "chain"}`):
    partial -> exchange(auto: hash on the group keys, or coalesce on rank 0 when every rank holds few rows) -> final ->
    exchange(gather to rank 0) -> tail operators (sort).  Nothing leaves HBM or the library between the stages; rank 0
    pulls the result, the other ranks pull an empty batch.  Same decisions as `final_aggregate` above."""
    tail = list(tail)
    head = [partial_spec, {"op": "exchange", "mode": "auto", "exprs": [{"col": c} for c in key_cols], "small_rows": small_rows}, final_spec]
    if tail and tail[0].get("op") == "sort":
        # SortExec per partition -> SortPreservingMergeExec on the root (test_tpch.plan.yaml:9-10): every rank sorts what it owns, the
        # sorted runs are gathered as they are and merged, instead of re-sorting the gathered rows
        s = tail[0]
        merge = {"op": "sort_preserving_merge", "keys": s["keys"], "fetch": s.get("fetch"), "runs": "batches"}
        return {"op": "chain", "ops": head + [s, {"op": "exchange", "mode": "gather", "root": 0, "keep_runs": True}, merge] + tail[1:]}
    return {"op": "chain", "ops": head + [{"op": "exchange", "mode": "gather", "root": 0}] + tail}
