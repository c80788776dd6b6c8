"""Pins the oracle (oracle/ops.py) and the data generator (datagen/) against the reference's own
golden vectors: python/pysail/tests/spark/__snapshots__/test_tpch.result.yaml, committed as
tests/golden/tpch_sf0001_result.json (script: tests/golden/make_golden.py)."""
import pytest

from oracle import ops, render
from sail_b200 import plans


def run_oracle(spec, *tables):
    out = ops.run_op(spec, *[ops.batch_from_arrow(t) for t in tables])
    return ops.batch_to_arrow(out)


@pytest.mark.parametrize("q", ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "q8", "q9", "q10", "q11", "q12", "q14", "q15", "q16", "q17", "q18", "q19", "q20", "q21", "q22"])
@pytest.mark.parametrize("strings", ["view", "utf8"])
def test_tpch_golden(q, strings, golden):
    from datagen import tpch
    tables = tpch.tables(0.001, strings=strings)
    st = "Utf8View" if strings == "view" else "Utf8"
    plan = plans.TPCH[q]() if q == "q6" else plans.TPCH[q](st)
    got = plans.execute(plan, tables, run_oracle)
    want = golden[q]
    assert got.schema.names == want["columns"]
    assert drop_unpinned(q, render.rows(got)) == drop_unpinned(q, want["rows"])


def drop_unpinned(q, rows):
    """Q10 returns c_comment: dbgen cuts comments out of a text pool built from a grammar this repository does not restate
    (datagen/tpch.py::text_pool), so that ONE column is left out of the comparison (the snapshot also trims its trailing blanks); every other column of every query is exact.
    (Q13 filters on comment text and is therefore not pinned at all: tests/test_gpu_relational.py checks it against the oracle.)"""
    return [[c.strip() for c in r[:-1]] for r in rows] if q == "q10" else rows      # (the snapshot table also trims an address that starts with a blank)


def test_c_pipelines_match_golden_and_python_oracle(golden):
    """oracle/cpipelines.c (the CPU-baseline port) against the golden Q1/Q6 snapshot and ops.py."""
    from datagen import tpch
    from oracle import cpipelines
    li = tpch.lineitem(0.001)
    rows = cpipelines.q1(li, plans.days("1998-09-24"), threads=3)
    want = golden["q1"]["rows"]
    assert len(rows) == len(want)

    def dec(v, s):
        sign = "-" if v < 0 else ""
        v = abs(v)
        return f"{sign}{v // 10**s}.{v % 10**s:0{s}d}"
    for r, w in zip(rows, want):
        rf, ls, sq, sp, sdp, sc, sd, cnt = r
        got = [rf, ls, dec(sq, 2), dec(sp, 2), dec(sdp, 4), dec(sc, 6),
               dec(sq * 10**4 // cnt, 6), dec(sp * 10**4 // cnt, 6), dec(sd * 10**4 // cnt, 6), str(cnt)]
        assert got == w
    n, s = cpipelines.q6(li, plans.days("1994-01-01"), plans.days("1995-01-01"), threads=2)
    assert dec(s, 4) == golden["q6"]["rows"][0][0]
