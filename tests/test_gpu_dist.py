"""Two ranks, two GPUs: partial aggregate -> hash repartition -> NCCL all-to-all (sailgpu_exchange) ->
final aggregate -> gather to root, against the single-process oracle.  Skipped with fewer than 2 GPUs."""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from datagen import tpch
        from sail_b200 import dist as sdist
        from sail_b200 import engine, plans
        from tests.util import assert_same, oracle_op

        ctx = engine.Context(rank)
        uid = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        backend = sdist.GpuBackend(ctx, rank, world)
        sf = 0.05
        first, n = sdist.shard_range(tpch.counts(sf)["orders"], rank, world)
        cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate", "l_shipmode"]
        shard = tpch.lineitem(sf, cols, first=first, n=n)
        # group by a long-string key too (l_shipmode has 7 values; exercises the view + heap exchange)
        C = plans.col
        names = shard.schema.names
        aggs = [("sum", C("l_extendedprice"), "s", "Decimal128(15,2)"), ("avg", C("l_discount"), "a", "Decimal128(15,2)"), ("count", None, "c", None)]
        scan = plans.scan("lineitem", names)
        partial_node = plans.aggregate(scan, "partial", ["l_returnflag", "l_shipmode"], aggs)
        final_node = plans.aggregate(partial_node, "final_partitioned", ["l_returnflag", "l_shipmode"], aggs)
        partial = backend.run(partial_node.spec, shard)
        mine = sdist.exchange_by_key(backend, partial, partial[0].schema, [0, 1])
        final = backend.run(final_node.spec, mine)
        root = sdist.gather_to_root(backend, final, final[0].schema)
        got = backend.to_host(root)
        if rank == 0:
            whole = tpch.lineitem(sf, cols)
            want = plans.execute(final_node, {"lineitem": whole}, oracle_op)
            assert_same(got, want)
        else:
            assert got.num_rows == 0
        # adaptive second phase: coalesce-on-root for few groups, hash exchange when forced (threshold 0)
        ident = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(final[0].schema.names)]}
        for small_rows, expect_on_root in ((1 << 14, True), (0, False)):
            partial2 = backend.run(partial_node.spec, shard)
            fin, on_root = sdist.final_aggregate(backend, partial2, partial2[0].schema, [0, 1], final_node.spec, small_rows=small_rows)
            assert on_root == expect_on_root
            if not on_root:
                fin = sdist.gather_to_root(backend, fin, fin[0].schema)
            if rank == 0:
                assert_same(backend.run_to_host(ident, fin), want)
        # the same plan as ONE chain inside the library (exchange operators), both exchange modes
        for small_rows in (1 << 14, 0):
            chain = sdist.two_phase_chain(partial_node.spec, final_node.spec, [0, 1], [ident], small_rows=small_rows)
            op = engine.GpuExec(chain, [shard.schema], ctx)
            op.push(shard)
            op.finish()
            res = op.collect()
            op.close()
            if rank == 0:
                assert_same(res, want)
            else:
                assert res.num_rows == 0
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_exchange():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_exchange_single_rank_is_a_move():
    import numpy as np
    import pyarrow as pa
    from sail_b200 import engine
    t = pa.table({"a": pa.array(np.arange(1000, dtype=np.int64)), "s": pa.array(["a long string value here"] * 1000, type=pa.string_view())})
    d = engine.to_device(t)
    out = engine.exchange([d], t.schema)
    ident = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(t.schema.names)]}
    op = engine.GpuExec(ident, [t.schema])
    op.push(out); op.finish()
    assert op.collect().equals(t)
