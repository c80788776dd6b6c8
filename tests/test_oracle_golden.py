"""Pins the oracle (oracle/ops.py) and the data generator (datagen/) against the reference's own
golden vectors: python/pysail/tests/spark/__snapshots__/test_tpch.result.yaml, committed as
tests/golden/tpch_sf0001_result.json (script: tests/golden/make_golden.py)."""
import pytest

from oracle import ops, render
from sail_b200 import plans


def run_oracle(spec, *tables):
    out = ops.run_op(spec, *[ops.batch_from_arrow(t) for t in tables])
    return ops.batch_to_arrow(out)


@pytest.mark.parametrize("q", ["q1", "q3", "q4", "q5", "q6", "q12"])
@pytest.mark.parametrize("strings", ["view", "utf8"])
def test_tpch_golden(q, strings, golden):
    from datagen import tpch
    tables = tpch.tables(0.001, strings=strings)
    st = "Utf8View" if strings == "view" else "Utf8"
    plan = plans.TPCH[q]() if q == "q6" else plans.TPCH[q](st)
    got = plans.execute(plan, tables, run_oracle)
    want = golden[q]
    assert got.schema.names == want["columns"]
    assert render.rows(got) == want["rows"]
