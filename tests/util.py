"""Shared helpers for the parity tests: run one operator spec / a whole plan through the oracle and
through the CUDA engine (C ABI) and compare the results as multisets of rendered rows."""
import pyarrow as pa

from oracle import ops, render


def oracle_op(spec, *tables):
    out = ops.run_op(spec, *[ops.batch_from_arrow(t) for t in tables])
    if isinstance(out, list):
        return [ops.batch_to_arrow(b) for b in out]
    return ops.batch_to_arrow(out)


def gpu_op(spec, *tables):
    from sail_b200 import engine
    return engine.run_op(spec, *tables)


def rows_of(t: pa.Table):
    return sorted(render.rows(t))


def assert_same(got: pa.Table, want: pa.Table, ordered: bool = False, float_cols=()):
    assert got.schema.names == want.schema.names, (got.schema.names, want.schema.names)
    assert [str(f.type) for f in got.schema] == [str(f.type) for f in want.schema], (got.schema, want.schema)
    g, w = render.rows(got), render.rows(want)
    if not ordered:
        g, w = sorted(g), sorted(w)
    assert len(g) == len(w), f"row count {len(g)} != {len(w)}"
    if not float_cols:
        assert g == w, first_diff(g, w)
        return
    for a, b in zip(g, w):
        for i, (x, y) in enumerate(zip(a, b)):
            if i in float_cols and x != "NULL" and y != "NULL":
                fx, fy = float(x), float(y)
                assert abs(fx - fy) <= 1e-6 * max(1.0, abs(fy)), (a, b)     # north_star: 1e-6 relative for float SUM/AVG
            else:
                assert x == y, (a, b)


def first_diff(g, w):
    for i, (a, b) in enumerate(zip(g, w)):
        if a != b:
            return f"first difference at row {i}:\n got  {a}\n want {b}"
    return "lengths differ"
