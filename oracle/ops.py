"""oracle/ops.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference's operator semantics.

This module is the *checker*: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import it.  The product path (sail_b200/) never does.

What it restates.  Sail's hot-path operators are DataFusion 53.1.0 operators over arrow 58.1.0
kernels (/root/reference/Cargo.toml:156-188, Cargo.lock:2021-2023) --
third-party crates that are NOT vendored under /root/reference.  Their published behaviour is
restated here and anchored on the reference's own call sites and golden
vectors:
  * FilterExec / batch_filter        crates/sail-physical-plan/src/streaming/filter.rs:104-116
  * ProjectionExec + arithmetic      crates/sail-plan/src/function/scalar/math.rs:48-181,580-583
  * AggregateExec (sum/avg/count/min/max, Partial/Final)
                                     crates/sail-plan/src/function/aggregate.rs:51-72,302-353,678-710
  * HashJoinExec                     crates/sail-execution/src/job_graph/planner.rs:137-147
  * SortExec / TopK                  crates/sail-session/src/planner.rs:7,34
  * Hash repartition                 crates/sail-execution/src/plan/shuffle_write.rs:173-196,226-232
  * Row round robin                  crates/sail-physical-plan/src/repartition.rs:46-84
Parity is PINNED: tests/test_oracle_golden.py runs the plans of
python/pysail/tests/spark/__snapshots__/test_tpch.plan.yaml through these functions on dbgen
SF0.001 data and reproduces test_tpch.result.yaml (fixtures under tests/golden/).

Representation: a column is (type string, numpy values, optional bool validity).  Decimal128 values
are Python ints (unscaled) in object arrays -- exact i128 arithmetic without a native type.
Operator specs are the same JSON objects the C ABI (include/sailgpu.h) accepts.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Optional

import numpy as np

INT_TYPES = {"Int8": np.int8, "Int16": np.int16, "Int32": np.int32, "Int64": np.int64,
             "UInt8": np.uint8, "UInt16": np.uint16, "UInt32": np.uint32, "UInt64": np.uint64}
FLOAT_TYPES = {"Float32": np.float32, "Float64": np.float64}
STRING_TYPES = ("Utf8", "Utf8View", "LargeUtf8")
# integer -> decimal coercion widths used by DataFusion's type coercion
INT_DECIMAL = {"Int8": (3, 0), "Int16": (5, 0), "Int32": (10, 0), "Int64": (20, 0),
               "UInt8": (3, 0), "UInt16": (5, 0), "UInt32": (10, 0), "UInt64": (20, 0)}
I128_MASK = (1 << 128) - 1


def wrap_i128(v: int) -> int:
    v &= I128_MASK
    return v - (1 << 128) if v >> 127 else v


def is_decimal(t: str) -> bool:
    return t.startswith("Decimal128")


def dec_ps(t: str):
    m = re.fullmatch(r"Decimal128\((\d+),\s*(-?\d+)\)", t)
    if not m:
        raise ValueError(f"bad decimal type {t}")
    return int(m.group(1)), int(m.group(2))


def dec_type(p: int, s: int) -> str:
    return f"Decimal128({p},{s})"


def is_string(t: str) -> bool:
    return t in STRING_TYPES


@dataclass
class Col:
    type: str
    data: np.ndarray
    valid: Optional[np.ndarray] = None   # None == all valid

    def __len__(self):
        return len(self.data)

    def validity(self) -> np.ndarray:
        return np.ones(len(self.data), dtype=bool) if self.valid is None else self.valid

    def take(self, idx: np.ndarray) -> "Col":
        return Col(self.type, self.data[idx], None if self.valid is None else self.valid[idx])


@dataclass
class Batch:
    names: list
    cols: list

    @property
    def num_rows(self):
        return len(self.cols[0]) if self.cols else 0

    def take(self, idx):
        return Batch(list(self.names), [c.take(idx) for c in self.cols])

    def select(self, proj):
        return Batch([self.names[i] for i in proj], [self.cols[i] for i in proj])


# ----------------------------------------------------------------------------------------------
# pyarrow bridge
# ----------------------------------------------------------------------------------------------
def type_from_arrow(t) -> str:
    import pyarrow as pa
    if pa.types.is_decimal128(t):
        return dec_type(t.precision, t.scale)
    table = {pa.bool_(): "Boolean", pa.int8(): "Int8", pa.int16(): "Int16", pa.int32(): "Int32",
             pa.int64(): "Int64", pa.uint8(): "UInt8", pa.uint16(): "UInt16", pa.uint32(): "UInt32",
             pa.uint64(): "UInt64", pa.float32(): "Float32", pa.float64(): "Float64",
             pa.date32(): "Date32", pa.string(): "Utf8", pa.string_view(): "Utf8View",
             pa.large_string(): "LargeUtf8"}
    if t in table:
        return table[t]
    raise ValueError(f"unsupported arrow type {t}")


def type_to_arrow(t: str):
    import pyarrow as pa
    if is_decimal(t):
        return pa.decimal128(*dec_ps(t))
    return {"Boolean": pa.bool_(), "Int8": pa.int8(), "Int16": pa.int16(), "Int32": pa.int32(),
            "Int64": pa.int64(), "UInt8": pa.uint8(), "UInt16": pa.uint16(), "UInt32": pa.uint32(),
            "UInt64": pa.uint64(), "Float32": pa.float32(), "Float64": pa.float64(),
            "Date32": pa.date32(), "Utf8": pa.string(), "Utf8View": pa.string_view(),
            "LargeUtf8": pa.large_string()}[t]


def col_from_arrow(arr) -> Col:
    import pyarrow as pa
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    t = type_from_arrow(arr.type)
    n = len(arr)
    valid = None
    if arr.null_count:
        valid = np.asarray(arr.is_valid())
    if is_decimal(t):
        raw = np.frombuffer(arr.buffers()[1], dtype=np.uint64,
                            count=2 * n, offset=arr.offset * 16).reshape(n, 2) if n else np.zeros((0, 2), np.uint64)
        data = np.empty(n, dtype=object)
        lo = raw[:, 0].tolist()
        hi = raw[:, 1].astype(np.int64).tolist()
        for i in range(n):
            data[i] = (hi[i] << 64) | lo[i]
    elif t == "Boolean":
        data = np.asarray(arr.fill_null(False)) if arr.null_count else np.asarray(arr)
        data = data.astype(bool)
    elif t == "Date32":
        data = np.asarray(arr.cast(pa.int32()).fill_null(0)) if arr.null_count else np.asarray(arr.cast(pa.int32()))
    elif is_string(t):
        data = np.empty(n, dtype=object)
        py = arr.to_pylist()
        for i in range(n):
            data[i] = b"" if py[i] is None else py[i].encode()
    else:
        data = np.asarray(arr.fill_null(0)) if arr.null_count else np.asarray(arr)
    return Col(t, data, valid)


def col_to_arrow(c: Col):
    import pyarrow as pa
    import decimal
    t = type_to_arrow(c.type)
    mask = None if c.valid is None else ~c.valid
    if is_decimal(c.type):
        p, s = dec_ps(c.type)
        ctx = decimal.Context(prec=60)
        vals = [None if (mask is not None and mask[i]) else
                ctx.scaleb(decimal.Decimal(int(c.data[i])), -s) for i in range(len(c.data))]
        return pa.array(vals, type=t)
    if is_string(c.type):
        vals = [None if (mask is not None and mask[i]) else c.data[i].decode() for i in range(len(c.data))]
        return pa.array(vals, type=t)
    if c.type == "Date32":
        return pa.array(np.asarray(c.data, dtype=np.int32), type=pa.int32(), mask=mask).cast(pa.date32())
    return pa.array(c.data, type=t, mask=mask)


def batch_from_arrow(tbl) -> Batch:
    import pyarrow as pa
    if isinstance(tbl, pa.RecordBatch):
        tbl = pa.Table.from_batches([tbl])
    return Batch(list(tbl.schema.names), [col_from_arrow(tbl.column(i)) for i in range(tbl.num_columns)])


def batch_to_arrow(b: Batch):
    import pyarrow as pa
    return pa.table([col_to_arrow(c) for c in b.cols], names=list(b.names))


# ----------------------------------------------------------------------------------------------
# expressions (DataFusion PhysicalExpr semantics)
# ----------------------------------------------------------------------------------------------
def _obj(vals) -> np.ndarray:
    a = np.empty(len(vals), dtype=object)
    for i, v in enumerate(vals):
        a[i] = v
    return a


def literal_col(lit, t: str, n: int) -> Col:
    if lit is None:
        if is_decimal(t) or is_string(t):
            data = _obj([0 if is_decimal(t) else b""] * n)
        elif t == "Boolean":
            data = np.zeros(n, dtype=bool)
        elif t == "Date32":
            data = np.zeros(n, dtype=np.int32)
        else:
            data = np.zeros(n, dtype=INT_TYPES.get(t) or FLOAT_TYPES[t])
        return Col(t, data, np.zeros(n, dtype=bool))
    if is_decimal(t):
        return Col(t, _obj([int(lit)] * n))
    if is_string(t):
        return Col(t, _obj([lit.encode() if isinstance(lit, str) else bytes(lit)] * n))
    if t == "Boolean":
        return Col(t, np.full(n, bool(lit)))
    if t == "Date32":
        return Col(t, np.full(n, int(lit), dtype=np.int32))
    if t in INT_TYPES:
        return Col(t, np.full(n, int(lit), dtype=INT_TYPES[t]))
    return Col(t, np.full(n, float(lit), dtype=FLOAT_TYPES[t]))


def _and_valid(a: Col, b: Col):
    if a.valid is None:
        return b.valid
    if b.valid is None:
        return a.valid
    return a.valid & b.valid


def _common_numeric(a: Col, b: Col):
    """Bring two operands to one arithmetic domain the way DataFusion's coercion would have
    (physical plans arrive already coerced; this only guards mixed int/decimal literals)."""
    ta, tb = a.type, b.type
    if is_decimal(ta) and tb in INT_DECIMAL:
        b = cast_col(b, dec_type(*INT_DECIMAL[tb]))
    elif is_decimal(tb) and ta in INT_DECIMAL:
        a = cast_col(a, dec_type(*INT_DECIMAL[ta]))
    elif ta in FLOAT_TYPES and tb not in FLOAT_TYPES:
        b = cast_col(b, ta)
    elif tb in FLOAT_TYPES and ta not in FLOAT_TYPES:
        a = cast_col(a, tb)
    elif ta in INT_TYPES and tb in INT_TYPES and ta != tb:
        wide = "Int64"
        a, b = cast_col(a, wide), cast_col(b, wide)
    return a, b


def decimal_result_type(op: str, p1, s1, p2, s2):
    """arrow-arith 58 decimal result types (SURVEY.md Appendix A 'Types & coercion')."""
    if op in ("+", "-"):
        s = max(s1, s2)
        return min(38, max(p1 - s1, p2 - s2) + s + 1), s
    if op == "*":
        return min(38, p1 + p2 + 1), s1 + s2
    if op == "/":
        s = min(38, s1 + 4)
        return min(38, p1 - s1 + s2 + s), s
    if op == "%":
        s = max(s1, s2)
        return min(38, min(p1 - s1, p2 - s2) + s), s
    raise ValueError(op)


def _trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _arith(op: str, a: Col, b: Col) -> Col:
    a, b = _common_numeric(a, b)
    valid = _and_valid(a, b)
    n = len(a)
    if is_decimal(a.type):
        p1, s1 = dec_ps(a.type)
        p2, s2 = dec_ps(b.type)
        p, s = decimal_result_type(op, p1, s1, p2, s2)
        ok = np.ones(n, bool) if valid is None else valid
        out = np.empty(n, dtype=object)
        if op in ("+", "-", "%"):
            lm, rm = 10 ** (s - s1), 10 ** (s - s2)
            for i in range(n):
                if not ok[i]:
                    out[i] = 0
                    continue
                x, y = a.data[i] * lm, b.data[i] * rm
                if op == "+":
                    out[i] = wrap_i128(x + y)
                elif op == "-":
                    out[i] = wrap_i128(x - y)
                else:
                    if y == 0:
                        raise ZeroDivisionError("Divide by zero")
                    out[i] = x - _trunc_div(x, y) * y
        elif op == "*":
            for i in range(n):
                out[i] = wrap_i128(a.data[i] * b.data[i]) if ok[i] else 0
        else:
            mul_pow = s - s1 + s2
            lm = 10 ** mul_pow if mul_pow > 0 else 1
            rm = 10 ** (-mul_pow) if mul_pow < 0 else 1
            for i in range(n):
                if not ok[i]:
                    out[i] = 0
                    continue
                if b.data[i] == 0:
                    raise ZeroDivisionError("Divide by zero")
                out[i] = _trunc_div(a.data[i] * lm, b.data[i] * rm)
        return Col(dec_type(p, s), out, valid)
    if a.type in FLOAT_TYPES:
        with np.errstate(all="ignore"):
            if op == "+":
                d = a.data + b.data
            elif op == "-":
                d = a.data - b.data
            elif op == "*":
                d = a.data * b.data
            elif op == "/":
                d = a.data / b.data
            else:
                d = np.fmod(a.data, b.data)
        return Col(a.type, d, valid)
    if a.type in INT_TYPES or a.type == "Date32":
        with np.errstate(all="ignore"):
            if op == "+":
                d = a.data + b.data          # numpy fixed-width ints wrap, like add_wrapping
            elif op == "-":
                d = a.data - b.data
            elif op == "*":
                d = a.data * b.data
            else:
                ok = np.ones(n, bool) if valid is None else valid
                if np.any((b.data == 0) & ok):
                    raise ZeroDivisionError("Divide by zero")
                safe = np.where(b.data == 0, 1, b.data)
                q = np.abs(a.data.astype(object)) // np.abs(safe.astype(object))
                q = np.where((a.data >= 0) == (safe >= 0), q, -q)
                d = q if op == "/" else a.data.astype(object) - q * safe.astype(object)
                d = np.asarray(d, dtype=a.data.dtype)
        return Col(a.type, d, valid)
    raise TypeError(f"arithmetic on {a.type}")


def _cmp_arrays(a: Col, b: Col):
    a, b = _common_numeric(a, b)
    if is_decimal(a.type):
        _, s1 = dec_ps(a.type)
        _, s2 = dec_ps(b.type)
        s = max(s1, s2)
        x = a.data if s == s1 else a.data * (10 ** (s - s1))
        y = b.data if s == s2 else b.data * (10 ** (s - s2))
        return x, y
    return a.data, b.data


def _compare(op: str, a: Col, b: Col) -> Col:
    x, y = _cmp_arrays(a, b)
    if op == "=":
        d = x == y
    elif op == "!=":
        d = x != y
    elif op == "<":
        d = x < y
    elif op == "<=":
        d = x <= y
    elif op == ">":
        d = x > y
    else:
        d = x >= y
    return Col("Boolean", np.asarray(d, dtype=bool), _and_valid(a, b))


def _kleene(op: str, a: Col, b: Col) -> Col:
    av, bv = a.validity(), b.validity()
    at, bt = a.data & av, b.data & bv            # definitely true
    af, bf = (~a.data) & av, (~b.data) & bv      # definitely false
    if op == "and":
        t, f = at & bt, af | bf
    else:
        t, f = at | bt, af & bf
    valid = t | f
    return Col("Boolean", t, None if valid.all() else valid)


def like_to_regex(pattern: str) -> "re.Pattern":
    out = []
    i = 0
    while i < len(pattern):
        ch = pattern[i]
        if ch == "\\" and i + 1 < len(pattern):
            out.append(re.escape(pattern[i + 1]))
            i += 2
            continue
        out.append(".*" if ch == "%" else "." if ch == "_" else re.escape(ch))
        i += 1
    return re.compile("".join(out).encode(), re.S)


def cast_col(c: Col, to: str) -> Col:
    if c.type == to:
        return c
    n = len(c)
    if is_decimal(to):
        p, s = dec_ps(to)
        if is_decimal(c.type):
            _, s0 = dec_ps(c.type)
            if s >= s0:
                m = 10 ** (s - s0)
                return Col(to, _obj([int(v) * m for v in c.data]), c.valid)
            d = 10 ** (s0 - s)
            # arrow rescale down rounds half away from zero
            def rs(v):
                q, r = divmod(abs(v), d)
                q += 1 if 2 * r >= d else 0
                return q if v >= 0 else -q
            return Col(to, _obj([rs(int(v)) for v in c.data]), c.valid)
        if c.type in INT_TYPES:
            m = 10 ** s
            return Col(to, _obj([int(v) * m for v in c.data]), c.valid)
        if c.type in FLOAT_TYPES:
            m = 10 ** s
            return Col(to, _obj([int(round(float(v) * m)) for v in c.data]), c.valid)
    if to in FLOAT_TYPES:
        if is_decimal(c.type):
            _, s0 = dec_ps(c.type)
            return Col(to, np.array([int(v) / (10 ** s0) for v in c.data], dtype=FLOAT_TYPES[to]), c.valid)
        return Col(to, c.data.astype(FLOAT_TYPES[to]), c.valid)
    if to in INT_TYPES:
        if is_decimal(c.type):
            _, s0 = dec_ps(c.type)
            d = 10 ** s0
            return Col(to, np.array([_trunc_div(int(v), d) for v in c.data], dtype=INT_TYPES[to]), c.valid)
        if c.type in FLOAT_TYPES:
            return Col(to, np.trunc(c.data).astype(INT_TYPES[to]), c.valid)
        return Col(to, c.data.astype(INT_TYPES[to]), c.valid)
    if to == "Date32" and c.type in INT_TYPES:
        return Col(to, c.data.astype(np.int32), c.valid)
    if is_string(to) and is_string(c.type):
        return Col(to, c.data, c.valid)
    raise TypeError(f"cast {c.type} -> {to}")


def civil_from_days(z: np.ndarray):
    """days since 1970-01-01 -> (year, month, day), proleptic Gregorian (Hinnant's algorithm)."""
    z = z.astype(np.int64) + 719468
    era = np.floor_divide(z, 146097)
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = np.where(mp < 10, mp + 3, mp - 9)
    y = np.where(m <= 2, y + 1, y)
    return y, m, d


def eval_expr(b: Batch, e: dict) -> Col:
    n = b.num_rows
    if "col" in e:
        return b.cols[e["col"]]
    if "lit" in e:
        return literal_col(e["lit"], e["type"], n)
    if "op" in e:
        op = e["op"]
        l, r = eval_expr(b, e["l"]), eval_expr(b, e["r"])
        if op in ("+", "-", "*", "/", "%"):
            return _arith(op, l, r)
        if op in ("=", "!=", "<", "<=", ">", ">="):
            return _compare(op, l, r)
        if op in ("and", "or"):
            return _kleene(op, l, r)
        raise ValueError(f"unknown binary op {op}")
    if "not" in e:
        c = eval_expr(b, e["not"])
        return Col("Boolean", ~c.data, c.valid)
    if "neg" in e:
        c = eval_expr(b, e["neg"])
        if is_decimal(c.type):
            return Col(c.type, _obj([wrap_i128(-int(v)) for v in c.data]), c.valid)
        return Col(c.type, -c.data, c.valid)
    if "is_null" in e:
        c = eval_expr(b, e["is_null"])
        return Col("Boolean", ~c.validity())
    if "is_not_null" in e:
        c = eval_expr(b, e["is_not_null"])
        return Col("Boolean", c.validity().copy())
    if "cast" in e:
        return cast_col(eval_expr(b, e["cast"]), e["to"])
    if "case" in e:
        # CASE WHEN c1 THEN v1 ... ELSE ve END : first true branch wins; no else => NULL
        branches = [(eval_expr(b, w), eval_expr(b, t)) for w, t in e["case"]]
        rtype = branches[0][1].type
        els = eval_expr(b, e["else"]) if e.get("else") is not None else literal_col(None, rtype, n)
        if is_decimal(rtype):
            # unify decimal branches to the widest (DataFusion coerces THEN/ELSE types at plan time)
            ps = [dec_ps(x.type) for _, x in branches] + ([dec_ps(els.type)] if is_decimal(els.type) else [])
            s = max(q[1] for q in ps)
            p = min(38, max(q[0] - q[1] for q in ps) + s)
            rtype = dec_type(p, s)
            branches = [(w, cast_col(t, rtype)) for w, t in branches]
            els = cast_col(els, rtype)
        data = els.data.copy()
        valid = els.validity().copy()
        decided = np.zeros(n, dtype=bool)
        for w, t in branches:
            hit = w.data & w.validity() & ~decided
            data[hit] = t.data[hit]
            valid[hit] = t.validity()[hit]
            decided |= hit
        return Col(rtype, data, None if valid.all() else valid)
    if "in" in e:
        c = eval_expr(b, e["in"])
        lits = e["set"]
        hit = np.zeros(n, dtype=bool)
        for lit in lits:
            hit |= _compare("=", c, literal_col(lit["lit"], lit["type"], n)).data
        if e.get("negated"):
            hit = ~hit
        return Col("Boolean", hit, c.valid)
    if "like" in e:
        c = eval_expr(b, e["like"])
        rx = like_to_regex(e["pattern"])
        hit = np.array([rx.fullmatch(v) is not None for v in c.data], dtype=bool) if n else np.zeros(0, bool)
        if e.get("negated"):
            hit = ~hit
        return Col("Boolean", hit, c.valid)
    if "fn" in e:
        fn = e["fn"]
        args = [eval_expr(b, a) for a in e["args"]]
        if fn == "date_part":
            part = e["part"].lower()
            y, m, d = civil_from_days(args[0].data)
            v = {"year": y, "month": m, "day": d}[part]
            return Col("Int32", v.astype(np.int32), args[0].valid)
        if fn == "substr":
            start, length = int(e["start"]), e.get("length")
            lo = max(start - 1, 0)
            out = _obj([v.decode()[lo: None if length is None else lo + int(length)].encode() for v in args[0].data])
            return Col(args[0].type, out, args[0].valid)
        raise ValueError(f"unknown scalar function {fn}")
    raise ValueError(f"bad expression {e}")


# ----------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------
def op_filter(b: Batch, spec: dict) -> Batch:
    """FilterExec: rows where predicate is TRUE (NULL => dropped), input order preserved, then the
    embedded projection=[...] column subset."""
    p = eval_expr(b, spec["predicate"])
    keep = np.nonzero(p.data & p.validity())[0]
    out = b.take(keep)
    proj = spec.get("projection")
    return out if proj is None else out.select(proj)


def op_projection(b: Batch, spec: dict) -> Batch:
    cols, names = [], []
    for item in spec["exprs"]:
        cols.append(eval_expr(b, item["expr"]))
        names.append(item["name"])
    return Batch(names, cols)


def _key_tuple_rows(cols, n):
    """Rows as hashable tuples; NULL is represented by None (all-NULL keys group together)."""
    lists = []
    for c in cols:
        vals = c.data.tolist()
        if c.valid is not None:
            v = c.valid.tolist()
            vals = [x if ok else None for x, ok in zip(vals, v)]
        lists.append(vals)
    return list(zip(*lists)) if lists else [()] * n


def agg_state_types(fn: str, in_type: Optional[str]):
    """(state column types, final type) per DataFusion UDAF."""
    if fn == "count":
        return ["Int64"], "Int64"
    if fn in ("min", "max"):
        return [in_type], in_type
    if fn == "sum":
        if is_decimal(in_type):
            p, s = dec_ps(in_type)
            t = dec_type(min(38, p + 10), s)
        elif in_type in FLOAT_TYPES:
            t = "Float64"
        elif in_type.startswith("UInt"):
            t = "UInt64"
        else:
            t = "Int64"
        return [t], t
    if fn == "avg":
        if is_decimal(in_type):
            p, s = dec_ps(in_type)
            return ["UInt64", dec_type(min(38, p + 10), s)], dec_type(min(38, p + 4), min(38, s + 4))
        return ["UInt64", "Float64"], "Float64"
    raise ValueError(fn)


def _finish_avg(total, count, in_type, final_type):
    if is_decimal(final_type):
        _, s_in = dec_ps(in_type)
        _, s_out = dec_ps(final_type)
        return _trunc_div(total * (10 ** (s_out - s_in)), count)     # DecimalAverager: truncating
    return float(total) / float(count)


def op_aggregate(b: Batch, spec: dict) -> Batch:
    """AggregateExec.  mode: 'single' | 'partial' (emit state columns) | 'final' /
    'final_partitioned' (merge state columns: input = group cols ++ state cols in aggregate order).
    Group output order = first-seen (DataFusion's is unspecified; callers compare as multisets)."""
    mode = spec.get("mode", "single")
    n = b.num_rows
    gcols = [eval_expr(b, g["expr"]) for g in spec["group_by"]]
    gnames = [g["name"] for g in spec["group_by"]]
    keys = _key_tuple_rows(gcols, n)
    group_of = {}
    gid = np.empty(n, dtype=np.int64)
    first_row = []
    for i, k in enumerate(keys):
        j = group_of.get(k)
        if j is None:
            j = len(group_of)
            group_of[k] = j
            first_row.append(i)
        gid[i] = j
    ng = len(group_of)
    if not spec["group_by"] and ng == 0:
        ng = 1          # global aggregate over empty input still yields one row
        first_row = []
    out_cols = [c.take(np.array(first_row, dtype=np.int64)) for c in gcols] if spec["group_by"] else []
    out_names = list(gnames)
    merging = mode in ("final", "final_partitioned")
    state_pos = len(spec["group_by"])        # cursor into input state columns when merging
    for a in spec["aggs"]:
        fn = a["fn"]
        if merging:
            in_type = a["input_type"]
            st_types, final_t = agg_state_types(fn, in_type)
            states = [b.cols[state_pos + k] for k in range(len(st_types))]
            state_pos += len(st_types)
        else:
            arg = eval_expr(b, a["args"][0]) if a.get("args") else None
            in_type = arg.type if arg is not None else None
            st_types, final_t = agg_state_types(fn, in_type)
        # accumulate ------------------------------------------------------------------------
        cnt = [0] * ng
        acc = [None] * ng
        if merging:
            if fn == "count":
                v, ok = states[0].data.tolist(), states[0].validity().tolist()
                for i in range(n):
                    if ok[i]:
                        cnt[gid[i]] += int(v[i])
            elif fn == "avg":
                cv, sv, ok = states[0].data.tolist(), states[1].data.tolist(), states[1].validity().tolist()
                for i in range(n):
                    g = gid[i]
                    cnt[g] += int(cv[i])
                    if ok[i]:
                        acc[g] = sv[i] if acc[g] is None else acc[g] + sv[i]
            else:
                v, ok = states[0].data.tolist(), states[0].validity().tolist()
                for i in range(n):
                    if not ok[i]:
                        continue
                    g = gid[i]
                    if acc[g] is None:
                        acc[g] = v[i]
                    elif fn == "sum":
                        acc[g] = acc[g] + v[i]
                    elif fn == "min":
                        acc[g] = min(acc[g], v[i])
                    else:
                        acc[g] = max(acc[g], v[i])
        else:
            if arg is None:                      # count(*) / count(1)
                for i in range(n):
                    cnt[gid[i]] += 1
            else:
                v, ok = arg.data.tolist(), arg.validity().tolist()
                for i in range(n):
                    if not ok[i]:
                        continue
                    g = gid[i]
                    cnt[g] += 1
                    if fn == "count":
                        continue
                    if acc[g] is None:
                        acc[g] = v[i]
                    elif fn in ("sum", "avg"):
                        acc[g] = acc[g] + v[i]
                    elif fn == "min":
                        acc[g] = min(acc[g], v[i])
                    else:
                        acc[g] = max(acc[g], v[i])
        # emit ----------------------------------------------------------------------------
        def mk(t, vals):
            valid = np.array([x is not None for x in vals], dtype=bool)
            if is_decimal(t):
                data = _obj([wrap_i128(int(x)) if x is not None else 0 for x in vals])
            elif is_string(t):
                data = _obj([x if x is not None else b"" for x in vals])
            elif t == "Date32":
                data = np.array([x if x is not None else 0 for x in vals], dtype=np.int32)
            elif t == "Boolean":
                data = np.array([bool(x) if x is not None else False for x in vals], dtype=bool)
            else:
                dt = INT_TYPES.get(t) or FLOAT_TYPES[t]
                data = np.array([x if x is not None else 0 for x in vals]).astype(dt) if vals else np.zeros(0, dt)
            return Col(t, data, None if valid.all() else valid)
        name = a["name"]
        if mode == "partial":
            if fn == "count":
                out_cols.append(mk("Int64", cnt)); out_names.append(f"{name}[count]")
            elif fn == "avg":
                out_cols.append(mk("UInt64", cnt)); out_names.append(f"{name}[count]")
                out_cols.append(mk(st_types[1], acc)); out_names.append(f"{name}[sum]")
            else:
                out_cols.append(mk(st_types[0], acc)); out_names.append(f"{name}[{fn}]")
        else:
            if fn == "count":
                out_cols.append(mk("Int64", cnt))
            elif fn == "avg":
                vals = [None if acc[g] is None or cnt[g] == 0 else _finish_avg(acc[g], cnt[g], in_type, final_t)
                        for g in range(ng)]
                out_cols.append(mk(final_t, vals))
            else:
                out_cols.append(mk(final_t, acc))
            out_names.append(name)
    return Batch(out_names, out_cols)


def op_hash_join(left: Batch, right: Batch, spec: dict) -> Batch:
    """HashJoinExec: build = LEFT child.  'on' = [[left_col, right_col], ...] column indices.
    NullEqualsNothing unless spec['null_equals_null'].  Optional residual filter evaluated over
    (left ++ right) candidate pairs.  Output = left ++ right (semi/anti: one side) then projection.
    Output order: probe order, build matches in build order (reference order is unspecified)."""
    jt = spec.get("join_type", "inner")
    on = spec["on"]
    nen = spec.get("null_equals_null", False)
    lk = _key_tuple_rows([left.cols[a] for a, _ in on], left.num_rows)
    rk = _key_tuple_rows([right.cols[c] for _, c in on], right.num_rows)
    table = {}
    for i, k in enumerate(lk):
        if not nen and any(x is None for x in k):
            continue
        table.setdefault(k, []).append(i)
    li, ri = [], []
    for j, k in enumerate(rk):
        if not nen and any(x is None for x in k):
            continue
        for i in table.get(k, ()):
            li.append(i); ri.append(j)
    li = np.array(li, dtype=np.int64); ri = np.array(ri, dtype=np.int64)
    if spec.get("filter") is not None and len(li):
        pair = Batch(left.names + right.names, [c.take(li) for c in left.cols] + [c.take(ri) for c in right.cols])
        p = eval_expr(pair, spec["filter"])
        keep = p.data & p.validity()
        li, ri = li[keep], ri[keep]
    nl, nr = left.num_rows, right.num_rows

    def null_side(bt: Batch, n):
        return [Col(c.type, literal_col(None, c.type, n).data, np.zeros(n, bool)) for c in bt.cols]
    if jt == "inner":
        cols = [c.take(li) for c in left.cols] + [c.take(ri) for c in right.cols]
        names = left.names + right.names
    elif jt in ("left", "right", "full"):
        lcols = [c.take(li) for c in left.cols]
        rcols = [c.take(ri) for c in right.cols]
        parts_l, parts_r = [lcols], [rcols]
        if jt in ("left", "full"):
            miss = np.setdiff1d(np.arange(nl), li)
            parts_l.append([c.take(miss) for c in left.cols]); parts_r.append(null_side(right, len(miss)))
        if jt in ("right", "full"):
            miss = np.setdiff1d(np.arange(nr), ri)
            parts_l.append(null_side(left, len(miss))); parts_r.append([c.take(miss) for c in right.cols])

        def cat(parts):
            out = []
            for k in range(len(parts[0])):
                data = np.concatenate([p[k].data for p in parts])
                valid = np.concatenate([p[k].validity() for p in parts])
                out.append(Col(parts[0][k].type, data, None if valid.all() else valid))
            return out
        cols = cat(parts_l) + cat(parts_r)
        names = left.names + right.names
    elif jt in ("left_semi", "left_anti"):
        hit = np.zeros(nl, bool); hit[li] = True
        idx = np.nonzero(hit if jt == "left_semi" else ~hit)[0]
        cols, names = [c.take(idx) for c in left.cols], list(left.names)
    elif jt in ("right_semi", "right_anti"):
        hit = np.zeros(nr, bool); hit[ri] = True
        idx = np.nonzero(hit if jt == "right_semi" else ~hit)[0]
        cols, names = [c.take(idx) for c in right.cols], list(right.names)
    else:
        raise ValueError(jt)
    out = Batch(names, cols)
    proj = spec.get("projection")
    return out if proj is None else out.select(proj)


def op_nested_loop_join(left: Batch, right: Batch, spec: dict) -> Batch:
    """NestedLoopJoinExec, inner (external; test_tpch.plan.yaml:333,661): every (left, right) pair that passes the filter;
    output = left ++ right then projection, build rows outermost."""
    assert spec.get("join_type", "inner") == "inner"
    nl, nr = left.num_rows, right.num_rows
    li = np.repeat(np.arange(nl, dtype=np.int64), nr)
    ri = np.tile(np.arange(nr, dtype=np.int64), nl)
    pair = Batch(left.names + right.names, [c.take(li) for c in left.cols] + [c.take(ri) for c in right.cols])
    if spec.get("filter") is not None and len(li):
        p = eval_expr(pair, spec["filter"])
        pair = pair.take(np.nonzero(p.data & p.validity())[0])
    proj = spec.get("projection")
    return pair if proj is None else pair.select(proj)


def sort_indices(b: Batch, keys) -> np.ndarray:
    """lexicographic, per-key asc/desc + nulls_first; stable (ties keep input order)."""
    import functools
    cols = [(eval_expr(b, k["expr"]), k.get("asc", True), k.get("nulls_first", k.get("asc", True))) for k in keys]
    rows = list(range(b.num_rows))
    vals = [(c.data.tolist(), c.validity().tolist(), asc, nf) for c, asc, nf in cols]

    def cmp(i, j):
        for data, ok, asc, nf in vals:
            a_ok, b_ok = ok[i], ok[j]
            if not a_ok or not b_ok:
                if a_ok == b_ok:
                    continue
                return (-1 if not a_ok else 1) * (1 if nf else -1)
            x, y = data[i], data[j]
            if x == y:
                continue
            r = -1 if x < y else 1
            return r if asc else -r
        return 0
    rows.sort(key=functools.cmp_to_key(cmp))
    return np.array(rows, dtype=np.int64)


def op_sort(b: Batch, spec: dict) -> Batch:
    """SortExec (+ TopK when 'fetch' is set)."""
    idx = sort_indices(b, spec["keys"])
    if spec.get("fetch") is not None:
        idx = idx[: int(spec["fetch"])]
    return b.take(idx)


def concat_batches(parts: list) -> Batch:
    first = parts[0]
    cols = []
    for i, c in enumerate(first.cols):
        data = np.concatenate([p.cols[i].data for p in parts])
        valid = None if all(p.cols[i].valid is None for p in parts) else np.concatenate([p.cols[i].validity() for p in parts])
        cols.append(Col(c.type, data, valid))
    return Batch(list(first.names), cols)


def op_sort_preserving_merge(runs: list, spec: dict) -> Batch:
    """SortPreservingMergeExec (external; test_tpch.plan.yaml:9-10): stable k-way merge of sorted runs, ties to the earlier run --
    which is what a stable sort of the runs laid end to end produces."""
    return op_sort(concat_batches([r for r in runs if r.num_rows] or runs[:1]), spec)


def hash_partition_ids(b: Batch, exprs, n_parts: int) -> np.ndarray:
    """Partition id per row for Hash(exprs, n).  The reference's hash function is an unobservable
    implementation detail (SURVEY.md Appendix A 'Exchange'); the property tests pin only that equal
    keys land in one partition and every row is delivered exactly once.  This oracle mirrors the
    GPU path's function (splitmix64-combined 64-bit key hash) so partition contents can also be
    compared exactly."""
    n = b.num_rows
    h = np.zeros(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for e in exprs:
            c = eval_expr(b, e)
            h = mix64(h ^ col_hash64(c))
    return (h % np.uint64(n_parts)).astype(np.int64)


def mix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def col_hash64(c: Col) -> np.ndarray:
    n = len(c)
    if is_decimal(c.type):
        lo = np.array([int(v) & 0xFFFFFFFFFFFFFFFF for v in c.data], dtype=np.uint64)
        hi = np.array([(int(v) >> 64) & 0xFFFFFFFFFFFFFFFF for v in c.data], dtype=np.uint64)
        # <=18-digit decimals are keyed on their low 64 bits (they fit), wider ones on all 128
        h = mix64(lo) if dec_ps(c.type)[0] <= 18 else mix64(lo ^ mix64(hi))
    elif is_string(c.type):
        out = np.zeros(n, dtype=np.uint64)
        for i, v in enumerate(c.data):
            x = np.uint64(len(v))
            for k in range(0, len(v), 8):
                x = mix64(np.array([x ^ np.uint64(int.from_bytes(v[k:k + 8], "little"))], dtype=np.uint64))[0]
            out[i] = x
        h = mix64(out)
    elif c.type in FLOAT_TYPES:
        h = mix64(c.data.astype(np.float64).view(np.uint64))
    else:
        h = mix64(c.data.astype(np.int64).view(np.uint64))
    if c.valid is not None:
        h = np.where(c.valid, h, np.uint64(0x6E756C6C6E756C6C))
    return h


def op_repartition(b: Batch, spec: dict):
    """RepartitionExec Hash(exprs, n) / ShuffleWriteExec partition step: list of n batches."""
    n_parts = int(spec["n"])
    if spec.get("scheme", "hash") == "round_robin_row":
        # RowRoundRobinPartitioner (crates/sail-physical-plan/src/repartition.rs:46-84)
        start = (int(spec.get("input_partition", 0)) * n_parts) // int(spec.get("num_input_partitions", 1))
        pid = (np.arange(b.num_rows, dtype=np.int64) + start) % n_parts
    else:
        pid = hash_partition_ids(b, spec["exprs"], n_parts)
    return [b.take(np.nonzero(pid == p)[0]) for p in range(n_parts)]


def run_op(spec: dict, *inputs: Batch):
    kind = spec["op"]
    if kind == "filter":
        return op_filter(inputs[0], spec)
    if kind == "projection":
        return op_projection(inputs[0], spec)
    if kind == "aggregate":
        return op_aggregate(inputs[0], spec)
    if kind == "hash_join":
        return op_hash_join(inputs[0], inputs[1], spec)
    if kind == "sort":
        return op_sort(inputs[0], spec)
    if kind == "repartition":
        return op_repartition(inputs[0], spec)
    if kind == "nested_loop_join":
        return op_nested_loop_join(inputs[0], inputs[1], spec)
    if kind == "sort_preserving_merge":
        return op_sort_preserving_merge(list(inputs), spec)
    raise ValueError(kind)
