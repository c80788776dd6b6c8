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


def assert_topk(got: pa.Table, full: pa.Table, keys, fetch=None, float_cols=()):
    """`got` = ORDER BY keys LIMIT fetch of the rows of `full` (all result rows, already ordered by `keys`), up to ties: the key
    columns agree position by position, and the rows `got` shows for one key value are among the rows `full` has for it (all of
    them, unless LIMIT cut the tie group).  Float64 columns are compared within 1e-6 relative, matched through the other cells."""
    import collections
    assert got.schema.names == full.schema.names, (got.schema.names, full.schema.names)
    assert [str(f.type) for f in got.schema] == [str(f.type) for f in full.schema], (got.schema, full.schema)
    g, w = render.rows(got), render.rows(full)
    n = len(w) if fetch is None else min(fetch, len(w))
    assert len(g) == n, f"row count {len(g)} != {n}"
    ki = [got.schema.names.index(k) for k in keys]
    exact = [i for i in range(got.num_columns) if i not in float_cols]

    def key(r):
        return tuple(r[i] for i in ki)

    def ident(r):
        return tuple(r[i] for i in exact)
    assert [key(r) for r in g] == [key(r) for r in w[:n]], first_diff([key(r) for r in g], [key(r) for r in w[:n]])
    have = collections.defaultdict(list)
    for r in w:
        have[key(r)].append(r)
    pools = {k: collections.Counter(ident(r) for r in rs) for k, rs in have.items()}
    floats = {ident(r): r for r in w} if float_cols else {}
    for r in g:
        pool = pools[key(r)]
        assert pool[ident(r)] > 0, f"row {r} is not a row of the full result (or shows up too often)"
        pool[ident(r)] -= 1
        if float_cols:
            ref = floats[ident(r)]
            for i in float_cols:
                if r[i] != "NULL" or ref[i] != "NULL":
                    assert abs(float(r[i]) - float(ref[i])) <= 1e-6 * max(1.0, abs(float(ref[i]))), (r, ref)
    # a tie group that lies wholly inside the first n rows is shown completely: sizes agree because the key sequences agree
