"""Regenerates tests/golden/tpch_sf0001_result.json from the reference's own snapshot
(python/pysail/tests/spark/__snapshots__/test_tpch.result.yaml).  Run in the build container only:
/root/reference does not exist on the GPU box, so the parsed vectors are committed.

    python tests/golden/make_golden.py
"""
import json
import os
import re

SRC = "/root/reference/python/pysail/tests/spark/__snapshots__/test_tpch.result.yaml"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tpch_sf0001_result.json")


def main():
    text = open(SRC).read()
    out = {}
    for block in text.split("\n---\n")[1:]:
        m = re.search(r'name: "test_derived_tpch_query_result\[(\d+)\](\.\d+)?"', block)
        if not m:
            continue
        # Q15 is three statements (create view / select / drop view): the select is snapshot [15].1
        if (m.group(2) or "") != (".1" if int(m.group(1)) == 15 else ""):
            continue
        lines = [ln.strip().strip("'") for ln in block.splitlines() if ln.strip().strip("'").startswith("| ")]
        rows = [[c.strip() for c in ln.strip("|").split("|")] for ln in lines]
        out[f"q{int(m.group(1))}"] = {"columns": rows[0], "rows": rows[1:]}
    with open(DST, "w") as f:
        json.dump({"source": "lakehq/sail python/pysail/tests/spark/__snapshots__/test_tpch.result.yaml "
                             "(DuckDB dbgen sf=0.001)", "queries": out}, f, indent=1)
    print("wrote", DST, sorted(out))


if __name__ == "__main__":
    main()
