"""The oracle against a second implementation.  oracle/ops.py restates DataFusion's operator semantics; where pyarrow's compute
kernels and Acero define the same thing (SURVEY.md section 8c "second opinion"), both are run on random tables with nulls,
duplicates and empty inputs and must agree.  Where they are known to differ from DataFusion (mean / sum of decimals, sort of
NaN, null ordering defaults) the expectation is computed in plain Python instead.  CPU only."""
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from sail_b200 import plans
from tests.util import assert_same, oracle_op


def table(n, seed, key_range=50, null_frac=0.15):
    rng = np.random.default_rng(seed)

    def m():
        return rng.random(n) < null_frac
    words = ["", "a", "ab", "Brand#14", "SMALL PLATED COPPER", "special requests", "дом", "x" * 40]
    return pa.table({
        "k": pa.array(rng.integers(0, key_range, n).astype(np.int64), mask=m()),
        "k2": pa.array(rng.integers(-3, 3, n).astype(np.int32), mask=m()),
        "v": pa.array(rng.integers(-10**6, 10**6, n).astype(np.int64), mask=m()),
        "f": pa.array(np.round(rng.normal(0, 100, n), 3), mask=m()),
        "d": pa.array([None if rng.random() < null_frac else decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**9, 10**9, n)], pa.decimal128(15, 2)),
        "s": pa.array([words[i] for i in rng.integers(0, len(words), n)], pa.string_view(), mask=m()),
        "day": pa.array(rng.integers(8000, 11000, n).astype(np.int32), mask=m()).cast(pa.date32()),
        "b": pa.array(rng.random(n) < 0.5, mask=m()),
    })


C = {n: i for i, n in enumerate(["k", "k2", "v", "f", "d", "s", "day", "b"])}


def col(n):
    return {"col": C[n]}


def utf8(t):
    """pyarrow's take / filter / group_by have no string_view kernels: run them on Utf8 ..."""
    return t.set_column(C["s"], "s", t["s"].cast(pa.string()))


def views(t):
    """... and cast the result back"""
    return t.set_column(t.schema.names.index("s"), "s", t["s"].cast(pa.string_view()))


def project(t, expr, name="x"):
    return oracle_op({"op": "projection", "exprs": [{"expr": expr, "name": name}]}, t).column(0)


@pytest.mark.parametrize("n", [0, 1, 500])
def test_comparisons_and_kleene_logic(n):
    t = table(n, 1)
    lt = plans.binop("<", col("v"), plans.lit(0, "Int64"))
    ge = plans.binop(">=", col("f"), plans.lit(10.0, "Float64"))
    assert project(t, lt).equals(pc.less(t["v"], 0))
    assert project(t, plans.binop("and", lt, ge)).equals(pc.and_kleene(pc.less(t["v"], 0), pc.greater_equal(t["f"], 10.0)))
    assert project(t, plans.binop("or", lt, col("b"))).equals(pc.or_kleene(pc.less(t["v"], 0), t["b"]))
    assert project(t, {"not": col("b")}).equals(pc.invert(t["b"]))
    assert project(t, {"is_null": col("s")}).equals(pc.is_null(t["s"]))
    # FilterExec keeps rows whose predicate is TRUE (NULL drops the row): pyarrow's filter with null_selection_behavior="drop"
    pred = plans.binop("or", lt, col("b"))
    got = oracle_op({"op": "filter", "predicate": pred, "projection": None}, t)
    want = views(utf8(t).filter(pc.or_kleene(pc.less(t["v"], 0), t["b"]), null_selection_behavior="drop"))
    assert_same(got, want, ordered=True)


@pytest.mark.parametrize("n", [0, 300])
def test_integer_float_and_date_expressions(n):
    t = table(n, 2)
    assert project(t, plans.binop("+", col("v"), col("k"))).equals(pc.add(t["v"], t["k"]))
    assert project(t, plans.binop("*", col("v"), plans.lit(3, "Int64"))).equals(pc.multiply(t["v"], 3))
    assert project(t, plans.binop("-", col("f"), plans.lit(0.5, "Float64"))).equals(pc.subtract(t["f"], 0.5))
    assert project(t, {"neg": col("v")}).equals(pc.negate(t["v"]))
    assert project(t, {"cast": col("k2"), "to": "Int64"}).equals(t["k2"].cast(pa.int64()))
    assert project(t, {"cast": col("v"), "to": "Float64"}).equals(t["v"].cast(pa.float64()))
    for part, fn in (("year", pc.year), ("month", pc.month), ("day", pc.day)):
        assert project(t, {"fn": "date_part", "part": part, "args": [col("day")]}).equals(fn(t["day"]).cast(pa.int32()))
    lo, hi = plans.date("1993-01-01"), plans.date("1996-06-30")
    between = plans.and_(plans.binop(">=", col("day"), lo), plans.binop("<=", col("day"), hi))
    import datetime
    want = pc.and_kleene(pc.greater_equal(t["day"], datetime.date(1993, 1, 1)), pc.less_equal(t["day"], datetime.date(1996, 6, 30)))
    assert project(t, between).equals(want)


def test_decimal_arithmetic_against_python_decimals():
    """arrow-arith result types (p, s) and exact values: checked with Python's Decimal (pyarrow's own decimal kernels use the
    same rules for + - *, which is asserted too)"""
    t = table(400, 3)
    d = t["d"]
    vals = d.to_pylist()
    mul = project(t, plans.binop("*", col("d"), col("d")))
    assert str(mul.type) == "decimal128(31, 4)" and mul.type == pc.multiply(d, d).type
    assert mul.to_pylist() == [None if v is None else v * v for v in vals]
    add = project(t, plans.binop("+", col("d"), plans.dec(100, 15, 2)))
    assert str(add.type) == "decimal128(16, 2)"
    assert add.to_pylist() == [None if v is None else v + 1 for v in vals]
    one_minus = project(t, plans.binop("-", plans.dec(1, 10, 0), col("d")))
    assert one_minus.to_pylist() == [None if v is None else 1 - v for v in vals]
    assert project(t, plans.binop("<", col("d"), plans.dec(0, 15, 2))).equals(pc.less(d, decimal.Decimal("0.00")))


@pytest.mark.parametrize("pattern", ["%requests%", "a%", "%x", "Brand#1_", "%L%P%", "", "%", "_", "special requests"])
def test_like_matches_pyarrow_match_like(pattern):
    t = table(300, 4)
    assert project(t, plans.like(col("s"), pattern)).equals(pc.match_like(t["s"].cast(pa.string()), pattern))
    assert project(t, plans.like(col("s"), pattern, True)).equals(pc.invert(pc.match_like(t["s"].cast(pa.string()), pattern)))


@pytest.mark.parametrize("start,length", [(1, 2), (1, None), (3, 5), (2, 0), (50, 3)])
def test_substr_counts_characters(start, length):
    t = table(300, 5)
    got = project(t, plans.substr(col("s"), start, length))
    stop = None if length is None else start - 1 + length
    want = pc.utf8_slice_codeunits(t["s"].cast(pa.string()), start - 1, stop)
    assert got.cast(pa.string()).equals(want)


def test_in_list_and_case():
    t = table(300, 6)
    got = project(t, {"in": col("k2"), "set": [plans.lit(v, "Int32") for v in (-1, 2)], "negated": False})
    assert got.to_pylist() == [None if v is None else v in (-1, 2) for v in t["k2"].to_pylist()]       # NULL IN (..) is NULL
    case = {"case": [[plans.binop("<", col("v"), plans.lit(0, "Int64")), plans.lit(-1, "Int64")], [plans.binop(">", col("v"), plans.lit(1000, "Int64")), col("k")]], "else": plans.lit(0, "Int64")}
    want = []
    for v, k in zip(t["v"].to_pylist(), t["k"].to_pylist()):
        want.append(-1 if (v is not None and v < 0) else (k if (v is not None and v > 1000) else 0))
    assert project(t, case).to_pylist() == want


@pytest.mark.parametrize("n", [0, 1, 2000])
@pytest.mark.parametrize("keys", [["k"], ["k2", "s"], ["b", "day"]])
def test_aggregate_against_acero_group_by(n, keys):
    """sum / min / max / count per group, NULL keys forming a group of their own; avg(int) is a Float64 mean"""
    t = table(n, 7)
    spec = {"op": "aggregate", "mode": "single", "group_by": [{"expr": col(k), "name": k} for k in keys],
            "aggs": [{"fn": "sum", "args": [col("v")], "name": "sv", "input_type": "Int64"}, {"fn": "min", "args": [col("v")], "name": "mn", "input_type": "Int64"},
                     {"fn": "max", "args": [col("f")], "name": "mx", "input_type": "Float64"}, {"fn": "count", "args": [col("v")], "name": "cv", "input_type": "Int64"},
                     {"fn": "count", "args": [], "name": "c", "input_type": None}, {"fn": "avg", "args": [col("v")], "name": "av", "input_type": "Int64"},
                     {"fn": "sum", "args": [col("f")], "name": "sf", "input_type": "Float64"}]}
    got = oracle_op(spec, t)
    u = utf8(t)
    ref = u.group_by(keys, use_threads=False).aggregate([("v", "sum"), ("v", "min"), ("f", "max"), ("v", "count"), ([], "count_all"), ("v", "mean"), ("f", "sum")])
    ref = ref.select(keys + ["v_sum", "v_min", "f_max", "v_count", "count_all", "v_mean", "f_sum"]).rename_columns(keys + ["sv", "mn", "mx", "cv", "c", "av", "sf"])
    ref = ref.set_column(len(keys) + 0, "sv", ref["sv"].cast(pa.int64()))
    if "s" in keys:
        ref = ref.set_column(keys.index("s"), "s", ref["s"].cast(pa.string_view()))
    assert_same(got, ref.cast(got.schema), float_cols=(len(keys) + 5, len(keys) + 6))


def test_two_phase_aggregate_equals_single():
    t = table(3000, 8)
    gb = ["k", "s"]
    aggs = [("sum", col("d"), "sd", "Decimal128(15,2)"), ("avg", col("d"), "ad", "Decimal128(15,2)"), ("avg", col("f"), "af", "Float64"), ("count", col("b"), "cb", "Boolean"),
            ("min", col("day"), "mn", "Date32"), ("max", col("d"), "mx", "Decimal128(15,2)")]
    src = plans.scan("t", list(C))
    single = plans.execute(plans.aggregate(src, "single", gb, aggs), {"t": t}, oracle_op)
    parts = [t.slice(0, 1000), t.slice(1000, 1500), t.slice(2500)]
    partial = plans.aggregate(src, "partial", gb, aggs)
    states = pa.concat_tables([plans.execute(partial, {"t": p}, oracle_op) for p in parts])
    final = oracle_op(plans.aggregate(partial, "final_partitioned", gb, aggs).spec, states)
    assert_same(final, single, float_cols=(4,))
    # avg over Decimal128(15,2): Decimal128(19,6), truncated toward zero (DataFusion's DecimalAverager), unlike pyarrow's mean
    by = {}
    for k, s, d in zip(t["k"].to_pylist(), t["s"].to_pylist(), t["d"].to_pylist()):
        if d is not None:
            by.setdefault((k, s), []).append(d)
    for k, s, ad in zip(single["k"].to_pylist(), single["s"].to_pylist(), single["ad"].to_pylist()):
        ds = by.get((k, s))
        if not ds:
            assert ad is None
            continue
        exact = sum(ds) / len(ds)
        assert ad == exact.quantize(decimal.Decimal("0.000001"), rounding=decimal.ROUND_DOWN)
    assert str(single.schema.field("ad").type) == "decimal128(19, 6)"


JOIN_TYPES = {"inner": "inner", "left": "left outer", "right": "right outer", "left_semi": "left semi", "left_anti": "left anti",
              "right_semi": "right semi", "right_anti": "right anti"}


@pytest.mark.parametrize("jt", list(JOIN_TYPES))
@pytest.mark.parametrize("nl,nr", [(0, 50), (50, 0), (200, 700)])
@pytest.mark.parametrize("on", [["k"], ["k", "k2"]])
def test_hash_join_against_acero(jt, nl, nr, on):
    """NullEqualsNothing: NULL keys never match (they still show up on the preserved side of outer / anti joins)"""
    left = table(nl, 9, key_range=40).select(["k", "k2", "v"]).rename_columns(["k", "k2", "lv"])
    right = table(nr, 10, key_range=40).select(["k", "k2", "f"]).rename_columns(["rk", "rk2", "rf"])
    spec = {"op": "hash_join", "join_type": jt, "on": [[["k", "k2"].index(c), ["k", "k2"].index(c)] for c in on], "filter": None, "projection": None}
    got = oracle_op(spec, left, right)
    want = left.join(right, keys=on, right_keys=["r" + c for c in on], join_type=JOIN_TYPES[jt], coalesce_keys=False, use_threads=False)
    assert_same(got, want.select(got.schema.names).cast(got.schema))


@pytest.mark.parametrize("keys", [[("v", True)], [("s", False), ("v", True)], [("f", False)], [("day", True), ("k2", False), ("d", True)], [("b", True), ("k", False)]])
@pytest.mark.parametrize("fetch", [None, 7])
def test_sort_against_pyarrow_sort_indices(keys, fetch):
    """Spark's default null ordering: ASC NULLS FIRST, DESC NULLS LAST == nulls are smaller than everything.  pyarrow places all
    nulls at one end whatever the direction, so the expectation sorts (is_valid, value) pairs per key instead."""
    t = table(800, 11)
    spec = {"op": "sort", "keys": [{"expr": col(k), "asc": asc, "nulls_first": asc} for k, asc in keys], "fetch": fetch}
    got = oracle_op(spec, t)
    sort_keys, u = [], t.set_column(C["s"], "s", t["s"].cast(pa.binary()))       # byte order == DataFusion's string order
    for i, (k, asc) in enumerate(keys):
        u = u.append_column(f"valid{i}", pc.is_valid(u[k]).cast(pa.int8()))
        sort_keys += [(f"valid{i}", "ascending" if asc else "descending"), (k, "ascending" if asc else "descending")]
    idx = pc.sort_indices(u, sort_keys=sort_keys)
    want = views(utf8(t).take(idx if fetch is None else idx[:fetch]))
    key_cols = [k for k, _ in keys]
    assert got.select(key_cols).equals(want.select(key_cols))          # ties may be ordered differently: keys must agree in order..
    if fetch is None:
        assert_same(got, want)                                         # ..and the rows as a multiset


def test_hash_repartition_is_a_partition_of_the_input():
    t = table(1000, 12)
    parts = oracle_op({"op": "repartition", "scheme": "hash", "exprs": [col("k"), col("s")], "n": 5}, t)
    assert len(parts) == 5 and sum(p.num_rows for p in parts) == t.num_rows
    assert_same(pa.concat_tables(parts), t)
    seen = {}
    for i, p in enumerate(parts):                                       # equal keys land in the same partition
        for key in set(zip(p["k"].to_pylist(), p["s"].to_pylist())):
            assert seen.setdefault(key, i) == i
