"""world_size-2 run (gloo, CPU) of the multi-GPU driver logic in sail_b200/dist.py: sharding, hash
repartition of partial-aggregate states, all-to-all, final aggregation on the owning rank, gather to
the root.  The operators themselves are executed by the oracle here (no GPU); on the GPU box the same
driver runs with GpuBackend (libsailgpu + NCCL) -- see tests/test_gpu_dist.py and bench.py --gpus N."""
import os
import socket

import pyarrow as pa
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from datagen import tpch
        from sail_b200 import dist as sdist
        from sail_b200 import plans
        from tests.util import oracle_op

        total_orders = tpch.counts(0.01)["orders"]
        first, n = sdist.shard_range(total_orders, rank, world)
        cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
        shard = tpch.lineitem(0.01, cols, first=first, n=n)
        sort_node = plans.q1()
        final_node = sort_node.inputs[0]
        partial_node = final_node.inputs[0]
        backend = sdist.HostBackend(oracle_op, rank, world)
        partial = plans.execute(partial_node, {"lineitem": shard}, oracle_op)
        mine = sdist.exchange_by_key(backend, partial, partial.schema, [0, 1])
        final = backend.run(final_node.spec, mine)
        # every group is owned by exactly one rank
        owners = [None] * world
        dist.all_gather_object(owners, [tuple(r[:2]) for r in map(lambda x: list(x.values()), final.to_pylist())])
        if rank == 0:
            flat = [g for o in owners for g in o]
            assert len(flat) == len(set(flat)), f"a group was finalised on two ranks: {owners}"
        root = sdist.gather_to_root(backend, final, final.schema)
        whole = None
        if rank == 0:
            result = backend.run(sort_node.spec, root)
            whole = plans.execute(sort_node, {"lineitem": tpch.lineitem(0.01, cols)}, oracle_op)
            assert result.equals(whole), (result.to_pylist(), whole.to_pylist())
        else:
            assert root.num_rows == 0
        # the adaptive second phase: few groups -> coalesced on the root (one exchange); threshold 0 -> hash exchange
        for small_rows, expect_on_root in ((1 << 14, True), (0, False)):
            fin, on_root = sdist.final_aggregate(backend, partial, partial.schema, [0, 1], final_node.spec, small_rows=small_rows)
            assert on_root == expect_on_root
            if not on_root:
                fin = sdist.gather_to_root(backend, fin, fin.schema)
            if rank == 0:
                assert backend.run_to_host(sort_node.spec, fin).equals(whole)
            elif on_root:
                assert fin is None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from sail_b200 import dist as sdist
    for total in (0, 1, 7, 150000):
        for world in (1, 2, 3, 8):
            pieces = [sdist.shard_range(total, r, world) for r in range(world)]
            assert sum(n for _, n in pieces) == total
            pos = 0
            for first, n in pieces:
                assert first == pos
                pos += n


def test_two_rank_partial_final_aggregate_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
