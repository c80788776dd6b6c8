"""CPU tests of the host side of libsailgpu (no GPU): spec parsing, DataFusion/arrow type inference and
error reporting through `sailgpu_spec_validate`, checked against the oracle's result types on every node of the
TPC-H plans whose results are pinned by the reference's golden snapshots."""
import pyarrow as pa
import pytest

from sail_b200 import engine, plans
from tests.util import oracle_op


def walk(node, tables, fn):
    if node.spec["op"] == "scan":
        return tables[node.spec["table"]].select(node.spec["columns"])
    ins = [walk(c, tables, fn) for c in node.inputs]
    out = oracle_op(node.spec, *ins)
    fn(node, ins, out)
    return out


@pytest.mark.parametrize("q", ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "q8", "q9", "q11", "q12", "q13", "q14", "q15", "q16", "q17", "q18", "q19", "q20", "q21", "q22"])
def test_output_schema_of_every_plan_node_matches_the_oracle(q, tpch_tiny):
    seen = []

    def check(node, ins, out):
        got = engine.validate(node.spec, [t.schema for t in ins])
        assert got.names == out.schema.names, (node.spec["op"], got.names, out.schema.names)
        assert [str(f.type) for f in got] == [str(f.type) for f in out.schema], (node.spec, got, out.schema)
        seen.append(node.spec["op"])

    walk(plans.TPCH[q](), tpch_tiny, check)
    assert seen


def test_wide_group_keys_take_the_sort_based_aggregate(tpch_tiny):
    """Q10 groups by seven columns: more than the hash table packs (6 keys / 64 key bytes).  The aggregate is still accepted at
    plan time -- it runs as sort-based grouping (WideAggOp) -- with the schema DataFusion gives it."""
    seen = []

    def check(node, ins, out):
        got = engine.validate(node.spec, [t.schema for t in ins])
        assert got.names == out.schema.names and [str(f.type) for f in got] == [str(f.type) for f in out.schema]
        seen.append(node.spec["op"])

    walk(plans.TPCH["q10"](), tpch_tiny, check)
    assert seen.count("aggregate") == 2


def test_decimal_type_rules():
    s = pa.schema([("a", pa.decimal128(15, 2)), ("b", pa.decimal128(16, 2)), ("i", pa.int32())])
    def ty(e):
        return str(engine.validate({"op": "projection", "exprs": [{"expr": e, "name": "x"}]}, [s]).field(0).type)
    A, B, I = {"col": 0}, {"col": 1}, {"col": 2}
    assert ty(plans.binop("*", A, B)) == "decimal128(32, 4)"        # arrow-arith: p1+p2+1, s1+s2
    assert ty(plans.binop("+", A, B)) == "decimal128(17, 2)"
    assert ty(plans.binop("-", plans.dec(1, 10, 0), A)) == "decimal128(16, 2)"   # Int32(1) coerced to Decimal128(10,0)
    assert ty(plans.binop("/", A, B)) == "decimal128(21, 6)"        # scale s1+4, precision p1-s1+s2+scale
    assert ty(plans.binop("+", A, I)) == "decimal128(16, 2)"
    agg = {"op": "aggregate", "mode": "single", "group_by": [], "aggs": [{"fn": "sum", "args": [A], "name": "s"}, {"fn": "avg", "args": [A], "name": "a"},
                                                                          {"fn": "count", "args": [], "name": "c"}, {"fn": "avg", "args": [I], "name": "ai"}]}
    out = engine.validate(agg, [s])
    assert [str(f.type) for f in out] == ["decimal128(25, 2)", "decimal128(19, 6)", "int64", "double"]


def test_unsupported_and_invalid_specs_are_rejected_at_plan_time():
    s = pa.schema([("a", pa.int64()), ("s", pa.string_view())])
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "projection", "exprs": [{"expr": {"fn": "regexp_replace", "args": [{"col": 1}]}, "name": "x"}]}, [s])
    assert e.value.code == 2 and "regexp_replace" in str(e.value)
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "filter", "predicate": {"col": 7}, "projection": None}, [s])
    assert e.value.code == 1
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "repartition", "scheme": "round_robin_batch", "n": 4}, [s])
    assert e.value.code == 2            # Hash and the row round-robin run on the GPU; RoundRobinBatch only re-labels batches
    assert engine.validate({"op": "repartition", "scheme": "round_robin_row", "n": 4}, [s]).names == ["a", "s"]
    with pytest.raises(engine.SailGpuError) as e:      # residual filters on outer joins: reported while planning, not after the build side was consumed
        engine.validate({"op": "hash_join", "join_type": "left", "on": [[0, 0]], "filter": {"op": "<", "l": {"col": 0}, "r": {"col": 2}}}, [s, s])
    assert e.value.code == 2
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "sort", "keys": [{"expr": {"op": "+", "l": {"col": 0}, "r": {"col": 0}}, "asc": True}]}, [s])
    assert e.value.code == 2
    with pytest.raises(engine.SailGpuError):
        engine.validate({"op": "hash_join", "join_type": "full", "on": [[0, 0]]}, [s, s])
    with pytest.raises(engine.SailGpuError):
        engine.validate({"op": "nonsense"}, [s])


def test_exchange_and_chain_specs_validate_without_a_gpu():
    """the shuffle boundary as an operator: schema passes through, key expressions are checked at plan time"""
    from sail_b200 import dist as sdist
    s = pa.schema([("k", pa.int64()), ("s", pa.string_view()), ("v", pa.decimal128(15, 2))])
    out = engine.validate({"op": "exchange", "mode": "auto", "exprs": [{"col": 0}, {"col": 1}]}, [s])
    assert out.names == s.names and [str(f.type) for f in out] == [str(f.type) for f in s]
    with pytest.raises(engine.SailGpuError):
        engine.validate({"op": "exchange", "mode": "hash", "exprs": [{"col": 7}]}, [s])
    with pytest.raises(engine.SailGpuError):
        engine.validate({"op": "exchange", "mode": "broadcast"}, [s])
    partial = {"op": "aggregate", "mode": "partial", "group_by": [{"expr": {"col": 0}, "name": "k"}],
               "aggs": [{"fn": "sum", "args": [{"col": 2}], "name": "sv", "input_type": "Decimal128(15,2)"}]}
    final = {"op": "aggregate", "mode": "final_partitioned", "group_by": [{"expr": {"col": 0}, "name": "k"}],
             "aggs": [{"fn": "sum", "name": "sv", "input_type": "Decimal128(15,2)"}]}
    out = engine.validate(sdist.two_phase_chain(partial, final, [0]), [s])
    assert out.names == ["k", "sv"] and str(out.field(1).type) == "decimal128(25, 2)"


def test_width_limits_of_one_operator_are_reported_at_plan_time():
    """a streaming pipeline writes at most 24 output columns and an aggregate carries at most 16 accumulators (vm.h): both are
    answered by sailgpu_spec_validate, not after the first batch arrived (ClickBench [29] met the first one at run time once)"""
    wide = pa.schema([(f"c{i}", pa.int64()) for i in range(30)])
    # picking / renaming columns passes buffers on: no limit
    assert len(engine.validate({"op": "projection", "exprs": [{"expr": {"col": i}, "name": f"x{i}"} for i in range(30)]}, [wide])) == 30
    computed = {"op": "projection", "exprs": [{"expr": plans.binop("+", {"col": i % 10}, plans.lit(i, "Int64")), "name": f"x{i}"} for i in range(25)]}
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate(computed, [wide])
    assert e.value.code == 2 and "24 output columns" in str(e.value)
    computed["exprs"] = computed["exprs"][:24]
    assert len(engine.validate(computed, [wide])) == 24
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "filter", "predicate": plans.binop(">", {"col": 0}, plans.lit(0, "Int64")), "projection": None}, [wide])
    assert e.value.code == 2            # 30 columns: more than 20 column buffers staged per tile
    half = pa.schema([(f"c{i}", pa.int64()) for i in range(13)])
    other = pa.schema([(f"d{i}", pa.int64()) for i in range(13)])
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "nested_loop_join", "join_type": "inner", "filter": None, "projection": None}, [half, other])
    assert e.value.code == 2 and "24 output columns" in str(e.value)
    assert len(engine.validate({"op": "nested_loop_join", "join_type": "inner", "filter": None, "projection": list(range(24))}, [half, other])) == 24
    sums = [{"fn": "sum", "args": [plans.binop("+", {"col": 0}, plans.lit(i, "Int64"))], "name": f"s{i}", "input_type": "Int64"} for i in range(17)]
    with pytest.raises(engine.SailGpuError) as e:
        engine.validate({"op": "aggregate", "mode": "single", "group_by": [], "aggs": sums}, [wide])
    assert e.value.code == 2 and "16" in str(e.value)
    assert len(engine.validate({"op": "aggregate", "mode": "single", "group_by": [], "aggs": sums[:16]}, [wide])) == 16
