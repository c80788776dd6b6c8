"""The 22 TPC-H plans of sail_b200/plans.py carry the same operators as the reference's physical plans
(python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml, condensed into tests/golden/tpch_plan_ops.json by
tests/golden/make_plan_golden.py): the same join types, the same aggregate modes, the same TopK fetches."""
import collections
import json
import os

import pytest

from sail_b200 import plans

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tpch_plan_ops.json")))["queries"]
JOIN_TYPE = {"inner": "Inner", "left": "Left", "right": "Right", "left_semi": "LeftSemi", "left_anti": "LeftAnti",
             "right_semi": "RightSemi", "right_anti": "RightAnti"}
AGG_MODE = {"partial": "Partial", "final": "Final", "final_partitioned": "FinalPartitioned", "single": "Single"}


def operators(node, acc):
    spec = node.spec
    if spec["op"] == "hash_join":
        acc["hash_joins"][JOIN_TYPE[spec["join_type"]]] += 1
    elif spec["op"] == "nested_loop_join":
        acc["nested_loop_joins"]["Inner"] += 1
    elif spec["op"] == "aggregate":
        acc["aggregates"][AGG_MODE[spec["mode"]]] += 1
    elif spec["op"] == "sort":
        acc["sorts"] += 1
        if spec.get("fetch") is not None:
            acc["topk"].append(spec["fetch"])
    for child in node.inputs:
        operators(child, acc)
    return acc


def test_every_query_of_the_snapshot_has_a_plan():
    assert sorted(GOLDEN) == sorted(plans.TPCH)


@pytest.mark.parametrize("query", sorted(GOLDEN, key=lambda q: int(q[1:])))
def test_plan_has_the_operators_of_the_reference_plan(query):
    acc = operators(plans.TPCH[query](), {"hash_joins": collections.Counter(), "nested_loop_joins": collections.Counter(),
                                          "aggregates": collections.Counter(), "topk": [], "sorts": 0})
    ref = GOLDEN[query]
    assert dict(acc["hash_joins"]) == ref["hash_joins"]
    assert dict(acc["nested_loop_joins"]) == ref["nested_loop_joins"]
    assert dict(acc["aggregates"]) == ref["aggregates"]
    assert acc["topk"] == ref["topk"]
    assert acc["sorts"] == ref["sorts"]
