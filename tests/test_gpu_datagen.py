"""The GPU data generator (datagen/tpch_dbgen_gpu.cu) against the C generator that reproduces the reference's golden
snapshot: every generated column bit-identical, at an offset into the key space and across a chunk boundary."""
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


def host_table(dev_batch):
    from sail_b200 import engine
    spec = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(dev_batch.schema.names)]}
    op = engine.GpuExec(spec, [dev_batch.schema])
    op.push(dev_batch.borrow())
    op.finish()
    t = op.collect()
    op.close()
    return t


@pytest.mark.parametrize("sf,first,n", [(0.01, 0, 15000), (1.0, 777_001, 40_003), (100.0, 149_000_000, 20_000)])
def test_gpu_generator_matches_the_c_generator(sf, first, n):
    from datagen import tpch, tpch_gpu
    ocols = ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority", "o_totalprice", "o_orderstatus"]
    lcols = list(tpch_gpu.LINEITEM_TYPES)
    o_dev, l_dev = tpch_gpu.generate(sf, first, n, ocols, lcols)
    o_want = tpch.orders(sf, ocols, first=first, n=n)
    l_want = tpch.lineitem(sf, lcols, first=first, n=n)
    o_got, l_got = host_table(o_dev), host_table(l_dev)
    assert o_got.num_rows == n and l_got.num_rows == l_want.num_rows
    for name in ocols:
        assert o_got.column(name).combine_chunks().equals(o_want.column(name).combine_chunks()), name
    for name in lcols:
        assert l_got.column(name).combine_chunks().equals(l_want.column(name).combine_chunks()), name
