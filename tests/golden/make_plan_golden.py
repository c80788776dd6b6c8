"""Regenerates tests/golden/tpch_plan_ops.json (and clickbench_plan_ops.json, see clickbench()) from the reference's plan snapshot
(python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml): per query, the operators that carry semantics -- join types of
HashJoinExec / NestedLoopJoinExec, modes of AggregateExec, TopK fetch of SortExec.  Run in the build container only.

    python tests/golden/make_plan_golden.py
"""
import collections
import json
import os
import re

SRC = "/root/reference/python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tpch_plan_ops.json")


def main():
    text = open(SRC).read()
    out = {}
    for block in text.split("\n---\n")[1:]:
        m = re.search(r'name: "test_derived_tpch_query_plan\[(\d+)\](\.\d+)?"', block)
        if not m:
            continue
        if (m.group(2) or "") != (".1" if int(m.group(1)) == 15 else ""):      # Q15: create view / select / drop view
            continue
        joins = collections.Counter(re.findall(r"HashJoinExec: mode=\w+, join_type=(\w+)", block))
        nlj = collections.Counter(re.findall(r"NestedLoopJoinExec: join_type=(\w+)", block))
        aggs = collections.Counter(re.findall(r"AggregateExec: mode=(\w+)", block))
        fetch = [int(x) for x in re.findall(r"SortExec: TopK\(fetch=(\d+)\)", block)]
        out[f"q{int(m.group(1))}"] = {"hash_joins": dict(joins), "nested_loop_joins": dict(nlj), "aggregates": dict(aggs), "topk": fetch,
                                     "sorts": len(re.findall(r"\bSortExec:", block))}
    with open(DST, "w") as f:
        json.dump({"source": "lakehq/sail python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml", "queries": out}, f, indent=1, sort_keys=True)
    print("wrote", DST, len(out))


def clickbench():
    """tests/golden/clickbench_plan_ops.json: per ClickBench query of the reference's plan snapshot (taken on an EMPTY hits table,
    python/pysail/tests/spark/test_clickbench.py:122-158) the pieces that do not depend on table size -- TopK fetch,
    GlobalLimitExec skip, and whether count(DISTINCT) is planned as two stacked aggregates (the inner one grouping by `alias1`)."""
    src = "/root/reference/python/pysail/tests/spark/__snapshots__/test_clickbench.plan.yaml"
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clickbench_plan_ops.json")
    out = {}
    for block in open(src).read().split("\n---\n"):
        m = re.search(r'name: "test_clickbench_query_plan\[(\d+)\]"', block)
        if not m:
            continue
        fetch = sorted({int(x) for x in re.findall(r"TopK\(fetch=(\d+)\)", block)})
        skip = re.findall(r"GlobalLimitExec: skip=(\d+), fetch=(\d+)", block)
        out[f"c{int(m.group(1))}"] = {"topk": fetch[0] if fetch else None, "skip": int(skip[0][0]) if skip else 0,
                                     "limit": int(skip[0][1]) if skip else None, "two_level_distinct": "alias1" in block}
    with open(dst, "w") as f:
        json.dump({"source": "lakehq/sail python/pysail/tests/spark/__snapshots__/test_clickbench.plan.yaml", "queries": out}, f, indent=1, sort_keys=True)
    print("wrote", dst, len(out))


if __name__ == "__main__":
    main()
    clickbench()
