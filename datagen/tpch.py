"""datagen/tpch.py -- synthetic TPC-H tables as pyarrow Tables (dbgen-exact for the generated columns).

Neither product nor oracle: fabricates inputs for tests and bench.py.  The C generator
(datagen/tpch_dbgen.c) is compiled on first use into datagen/_build/ (built by
__graft_entry__.build() so it travels to the GPU box prebuilt).

Arrow layout mirrors what Sail hands its operators for Parquet-backed TPC-H tables (SURVEY.md
section 8a): keys Int64, money Decimal128(15,2), dates Date32, strings Utf8View (Sail reads Parquet
strings as views: crates/sail-common/src/config/application.yaml:375-381) or Utf8
(`strings="utf8"`, what `spark.createDataFrame(pandas)` yields in test_tpch.py:19-24).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtpch_dbgen.so")
_lib = None

MKTSEGMENT = ["AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"]
ORDERPRIORITY = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
SHIPINSTRUCT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
SHIPMODE = ["REG AIR", "AIR", "RAIL", "TRUCK", "MAIL", "FOB", "SHIP"]   # codes 5,6 pinned by golden Q12
NATIONS = [("ALGERIA", 0), ("ARGENTINA", 1), ("BRAZIL", 1), ("CANADA", 1), ("EGYPT", 4),
           ("ETHIOPIA", 0), ("FRANCE", 3), ("GERMANY", 3), ("INDIA", 2), ("INDONESIA", 2),
           ("IRAN", 4), ("IRAQ", 4), ("JAPAN", 2), ("JORDAN", 4), ("KENYA", 0), ("MOROCCO", 0),
           ("MOZAMBIQUE", 0), ("PERU", 1), ("CHINA", 2), ("ROMANIA", 3), ("SAUDI ARABIA", 4),
           ("VIETNAM", 2), ("RUSSIA", 3), ("UNITED KINGDOM", 3), ("UNITED STATES", 1)]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tpch_dbgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.tpch_gen_orders.restype = ctypes.c_int64
    return _lib


def counts(sf: float) -> dict:
    out = (ctypes.c_int64 * 4)()
    lib().tpch_counts_get(ctypes.c_double(sf), out)
    return {"part": out[0], "supplier": out[1], "customer": out[2], "orders": out[3]}


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ---- numpy -> Arrow column builders (zero Python loops over rows) --------------------------------
def decimal_from_int64(v: np.ndarray, precision: int = 15, scale: int = 2) -> pa.Array:
    n = len(v)
    raw = np.empty((n, 2), dtype=np.int64)
    raw[:, 0] = v
    raw[:, 1] = v >> 63
    return pa.Array.from_buffers(pa.decimal128(precision, scale), n, [None, pa.py_buffer(raw)])


def date32(v: np.ndarray) -> pa.Array:
    return pa.Array.from_buffers(pa.date32(), len(v), [None, pa.py_buffer(np.ascontiguousarray(v, dtype=np.int32))])


def strings_from_codes(codes: np.ndarray, values: list, kind: str = "view") -> pa.Array:
    """Dictionary codes -> Utf8View (views share one data buffer; <=12-byte strings inline) or Utf8."""
    n = len(codes)
    enc = [s.encode() for s in values]
    if kind == "utf8":
        lens = np.array([len(s) for s in enc], dtype=np.int64)
        offsets = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(lens[codes], out=offsets[1:])
        width = int(lens.max())
        mat = np.zeros((len(enc), width), dtype=np.uint8)
        for i, s in enumerate(enc):
            mat[i, : len(s)] = np.frombuffer(s, dtype=np.uint8)
        rows = mat[codes]
        mask = np.arange(width)[None, :] < lens[codes][:, None]
        data = rows[mask]
        return pa.Array.from_buffers(pa.string(), n, [None, pa.py_buffer(offsets), pa.py_buffer(data.tobytes())])
    heap = b"".join(enc)
    table = np.zeros((len(enc), 16), dtype=np.uint8)
    off = 0
    for i, s in enumerate(enc):
        table[i, 0:4] = np.frombuffer(np.int32(len(s)).tobytes(), dtype=np.uint8)
        if len(s) <= 12:
            table[i, 4:4 + len(s)] = np.frombuffer(s, dtype=np.uint8)
        else:
            table[i, 4:8] = np.frombuffer(s[:4], dtype=np.uint8)
            table[i, 8:12] = np.frombuffer(np.int32(0).tobytes(), dtype=np.uint8)
            table[i, 12:16] = np.frombuffer(np.int32(off).tobytes(), dtype=np.uint8)
        off += len(s)
    views = np.ascontiguousarray(table[codes])
    bufs = [None, pa.py_buffer(views)]
    if any(len(s) > 12 for s in enc):
        bufs.append(pa.py_buffer(heap))
    return pa.Array.from_buffers(pa.string_view(), n, bufs)


def chars(v: np.ndarray, kind: str = "view") -> pa.Array:
    """single-byte flag column (l_returnflag, l_linestatus, o_orderstatus) -> string array"""
    n = len(v)
    if kind == "utf8":
        offsets = np.arange(n + 1, dtype=np.int32)
        return pa.Array.from_buffers(pa.string(), n, [None, pa.py_buffer(offsets), pa.py_buffer(np.ascontiguousarray(v))])
    views = np.zeros((n, 16), dtype=np.uint8)
    views[:, 0] = 1
    views[:, 4] = v
    return pa.Array.from_buffers(pa.string_view(), n, [None, pa.py_buffer(views)])


# ---- tables --------------------------------------------------------------------------------------
LINEITEM_ALL = ["l_orderkey", "l_partkey", "l_suppkey", "l_linenumber", "l_quantity", "l_extendedprice",
                "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate", "l_commitdate",
                "l_receiptdate", "l_shipinstruct", "l_shipmode"]
ORDERS_ALL = ["o_orderkey", "o_custkey", "o_orderstatus", "o_totalprice", "o_orderdate",
              "o_orderpriority", "o_clerk", "o_shippriority"]


def _gen_orders_chunk(sf, first, n, want_o, want_l):
    dt = {"o_orderkey": "i8", "o_custkey": "i8", "o_orderstatus": "u1", "o_totalprice": "i8", "o_orderdate": "i4",
          "o_orderpriority": "u1", "o_clerk": "i8", "o_shippriority": "i4",
          "l_orderkey": "i8", "l_partkey": "i8", "l_suppkey": "i8", "l_linenumber": "i4", "l_quantity": "i8",
          "l_extendedprice": "i8", "l_discount": "i8", "l_tax": "i8", "l_returnflag": "u1", "l_linestatus": "u1",
          "l_shipdate": "i4", "l_commitdate": "i4", "l_receiptdate": "i4", "l_shipinstruct": "u1", "l_shipmode": "u1"}
    arrs = {}
    for name in ORDERS_ALL:
        arrs[name] = np.empty(n, dtype=dt[name]) if name in want_o else None
    for name in LINEITEM_ALL:
        arrs[name] = np.empty(7 * n, dtype=dt[name]) if name in want_l else None
    nl = lib().tpch_gen_orders(ctypes.c_double(sf), ctypes.c_int64(first), ctypes.c_int64(n),
                               *[_p(arrs[k]) for k in ORDERS_ALL + LINEITEM_ALL])
    for name in LINEITEM_ALL:
        if arrs[name] is not None:
            arrs[name] = arrs[name][:nl]
    return arrs


def gen_orders_lineitem_numpy(sf: float, orders_cols=(), lineitem_cols=(), first: int = 0, n: int | None = None,
                              threads: int | None = None) -> dict:
    """Raw numpy columns (money in int64 cents, dates as int32 days, flags as bytes, categoricals as codes)."""
    total = counts(sf)["orders"]
    if n is None:
        n = total - first
    threads = threads or min(32, os.cpu_count() or 1)
    chunk = max(1, min(n, max(100_000, (n + threads - 1) // threads)))
    jobs = [(first + s, min(chunk, n - s)) for s in range(0, n, chunk)]
    if len(jobs) == 1:
        parts = [_gen_orders_chunk(sf, jobs[0][0], jobs[0][1], set(orders_cols), set(lineitem_cols))]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda j: _gen_orders_chunk(sf, j[0], j[1], set(orders_cols), set(lineitem_cols)), jobs))
    out = {}
    for name in list(orders_cols) + list(lineitem_cols):
        out[name] = np.concatenate([p[name] for p in parts]) if len(parts) > 1 else parts[0][name]
    return out


def _to_arrow(name: str, v: np.ndarray, strings: str) -> pa.Array:
    if name in ("l_quantity", "l_extendedprice", "o_totalprice", "c_acctbal", "s_acctbal", "p_retailprice"):
        return decimal_from_int64(v)
    if name in ("l_discount", "l_tax"):
        return decimal_from_int64(v)          # already hundredths: 0.05 == 5
    if name.endswith("date"):
        return date32(v)
    if name in ("l_returnflag", "l_linestatus", "o_orderstatus"):
        return chars(v, strings)
    if name == "o_orderpriority":
        return strings_from_codes(v, ORDERPRIORITY, strings)
    if name == "l_shipinstruct":
        return strings_from_codes(v, SHIPINSTRUCT, strings)
    if name == "l_shipmode":
        return strings_from_codes(v, SHIPMODE, strings)
    if name == "c_mktsegment":
        return strings_from_codes(v, MKTSEGMENT, strings)
    if v.dtype == np.int32:
        return pa.array(v, type=pa.int32())
    return pa.array(v, type=pa.int64())


def lineitem(sf: float, columns=None, strings: str = "view", **kw) -> pa.Table:
    columns = list(columns or LINEITEM_ALL)
    raw = gen_orders_lineitem_numpy(sf, (), columns, **kw)
    return pa.table([_to_arrow(c, raw[c], strings) for c in columns], names=columns)


def orders(sf: float, columns=None, strings: str = "view", **kw) -> pa.Table:
    columns = list(columns or ORDERS_ALL)
    gen = [c for c in columns if c != "o_comment"]
    raw = gen_orders_lineitem_numpy(sf, gen or ["o_orderkey"], (), **kw)
    n_rows = len(raw[(gen or ["o_orderkey"])[0]])
    return pa.table([comments(c, n_rows, strings, kw.get("first", 0)) if c == "o_comment" else _to_arrow(c, raw[c], strings) for c in columns], names=columns)


def _phones(which: int, nationkey: np.ndarray, strings: str) -> pa.Array:
    n = len(nationkey)
    raw = np.empty(15 * n, dtype=np.uint8)
    lib().tpch_gen_phone(ctypes.c_int(which), ctypes.c_int64(0), ctypes.c_int64(n), _p(np.ascontiguousarray(nationkey, dtype=np.int64)), _p(raw))
    txt = raw.reshape(n, 15).view("S15").ravel()
    return pa.array([t.decode() for t in txt], type=pa.string_view() if strings == "view" else pa.string())


# dists.dss "colors" (p_name draws five of them per part)
COLORS = ("almond antique aquamarine azure beige bisque black blanched blue blush brown burlywood burnished chartreuse chiffon chocolate "
          "coral cornflower cornsilk cream cyan dark deep dim dodger drab firebrick floral forest frosted gainsboro ghost goldenrod green "
          "grey honeydew hot indian ivory khaki lace lavender lawn lemon light lime linen magenta maroon medium metallic midnight mint "
          "misty moccasin navajo navy olive orange orchid pale papaya peach peru pink plum powder puff purple red rose rosy royal saddle "
          "salmon sandy seashell sienna sky slate smoke snow spring steel tan thistle tomato turquoise violet wheat white yellow").split()
assert len(COLORS) == 92

_NOUNS = ("foxes ideas theodolites pinto beans instructions dependencies excuses platelets asymptotes courts dolphins multipliers sauternes "
          "warthogs frets dinos attainments somas patterns forges braids frays warhorses dugouts notornis epitaphs pearls tithes waters orbits "
          "gifts sheaves depths sentiments decoys realms pains grouches escapades packages requests accounts deposits").split()
_VERBS = ("sleep wake are cajole haggle nag use boost affix detect integrate maintain nod was lose sublate solve thrash promise engage "
          "hinder print x-ray breach eat grow impress mold poach serve run dazzle snooze doze unwind kindle play hang believe doubt").split()
_ADJS = ("furious sly careful blithe quick fluffy slow quiet ruthless thin close dogged daring brave stealthy permanent enticing idle busy "
         "regular final ironic even bold silent special pending unusual express").split()
_ADVS = ("sometimes always never furiously slyly carefully blithely quickly fluffily slowly quietly ruthlessly thinly closely doggedly "
         "daringly bravely stealthily permanently enticingly idly busily regularly finally ironically evenly boldly silently").split()
_PREPS = ("about above according to across after against along alongside of among around at atop before behind beneath beside besides "
          "between beyond by despite during except for from in place of inside instead of into near of on outside over past since through "
          "throughout to toward under until up upon without with within").split()
_POOL = None
TEXT_POOL_BYTES = 1 << 20


def text_pool() -> bytes:
    """The comment pool.  dbgen pre-generates 300 MB of sentences from the grammar in dists.dss; that grammar is not restated
    here, so this pool (1 MiB of adverb/adjective/noun/verb sentences over dbgen's word lists, fixed seed) is NOT dbgen's: the
    comment columns are realistic (`special requests`, `express packages` ... occur) but do not reproduce dbgen's bytes, and
    the three golden snapshots that depend on comment text (q2's output is empty; q10's c_comment column; q13's filter) are
    pinned only as far as they do not."""
    global _POOL
    if _POOL is None:
        rng = np.random.RandomState(933588178 % (2 ** 31))
        out, size = [], 0
        while size < TEXT_POOL_BYTES + 256:
            k = rng.randint(0, 4)
            w = lambda L: L[rng.randint(0, len(L))]
            if k == 0:
                sent = f"{w(_ADVS)} {w(_ADJS)} {w(_NOUNS)} {w(_VERBS)} {w(_PREPS)} the {w(_ADJS)} {w(_NOUNS)}"
            elif k == 1:
                sent = f"{w(_ADJS)} {w(_NOUNS)} {w(_VERBS)} {w(_ADVS)}"
            elif k == 2:
                sent = f"{w(_NOUNS)} {w(_VERBS)} {w(_ADVS)} {w(_PREPS)} the {w(_ADVS)} {w(_ADJS)} {w(_NOUNS)}"
            else:
                sent = f"{w(_ADJS)}, {w(_ADJS)} {w(_NOUNS)} {w(_PREPS)} the {w(_NOUNS)} {w(_VERBS)}"
            sent += ".;:?!"[rng.randint(0, 5)] if rng.randint(0, 8) == 0 else "."
            out.append(sent + " ")
            size += len(sent) + 1
        _POOL = "".join(out).encode()[:TEXT_POOL_BYTES]
    return _POOL


_TEXT_STREAMS = {"o_comment": (12, 49), "c_comment": (31, 73), "s_comment": (36, 63)}   # Seed[] index, average length


def comments(column: str, n: int, strings: str = "view", first: int = 0) -> pa.Array:
    """dbgen's TEXT(): offset and length draws on the column's own stream, cut out of text_pool()"""
    stream, avg = _TEXT_STREAMS[column]
    lo, hi = int(avg * 0.4), int(avg * 1.6)
    off, ln = np.empty(n, "i8"), np.empty(n, "i4")
    pool = text_pool()
    lib().tpch_gen_text(ctypes.c_int(stream), ctypes.c_int64(first), ctypes.c_int64(n), ctypes.c_int64(lo), ctypes.c_int64(hi),
                        ctypes.c_int64(len(pool)), _p(off), _p(ln))
    vals = [pool[o:o + l].decode() for o, l in zip(off.tolist(), ln.tolist())]
    if column == "s_comment":       # mk_supp: "Customer <noise>Complaints|Recommends" overwrites part of a few comments
        kind = np.empty(n, "u1")
        lib().tpch_gen_bbb(ctypes.c_int64(first), ctypes.c_int64(n), _p(kind))
        for i in np.nonzero(kind)[0].tolist():
            word = "Complaints" if kind[i] == 1 else "Recommends"
            t = vals[i]
            if len(t) < 25:
                t = t + " " * (25 - len(t))
            vals[i] = t[:3] + "Customer " + t[12:len(t) - 10] + word
    return pa.array(vals, type=pa.string_view() if strings == "view" else pa.string())


def addresses(which: int, n: int, strings: str = "view", first: int = 0) -> pa.Array:
    ln, raw = np.empty(n, "i4"), np.zeros(40 * n, dtype=np.uint8)
    lib().tpch_gen_address(ctypes.c_int(which), ctypes.c_int64(first), ctypes.c_int64(n), _p(ln), _p(raw))
    rows = raw.reshape(n, 40)
    vals = [bytes(rows[i, :l]).decode() for i, l in enumerate(ln.tolist())]
    return pa.array(vals, type=pa.string_view() if strings == "view" else pa.string())


def part_names(n: int, strings: str = "view", first: int = 0) -> pa.Array:
    idx = np.empty(5 * n, dtype=np.uint8)
    lib().tpch_gen_pname(ctypes.c_int64(first), ctypes.c_int64(n), ctypes.c_int(1), _p(idx))
    vals = [" ".join(COLORS[j] for j in row) for row in idx.reshape(n, 5).tolist()]
    return pa.array(vals, type=pa.string_view() if strings == "view" else pa.string())


def customer(sf: float, columns=None, strings: str = "view") -> pa.Table:
    n = counts(sf)["customer"]
    a = {"c_custkey": np.empty(n, "i8"), "c_nationkey": np.empty(n, "i8"), "c_acctbal": np.empty(n, "i8"),
         "c_mktsegment": np.empty(n, "u1")}
    lib().tpch_gen_customer(ctypes.c_double(sf), ctypes.c_int64(0), ctypes.c_int64(n),
                            _p(a["c_custkey"]), _p(a["c_nationkey"]), _p(a["c_acctbal"]), _p(a["c_mktsegment"]))
    columns = list(columns or list(a.keys()) + ["c_name", "c_phone"])

    def one(c):
        if c == "c_name":       # dbgen: "Customer#%09d" (18 bytes: always a long view)
            txt = np.char.add("Customer#", np.char.zfill(a["c_custkey"].astype(str), 9))
            return pa.array(txt.tolist(), type=pa.string_view() if strings == "view" else pa.string())
        if c == "c_phone":
            return _phones(0, a["c_nationkey"], strings)
        if c == "c_address":
            return addresses(0, n, strings)
        if c == "c_comment":
            return comments(c, n, strings)
        return _to_arrow(c, a[c], strings)
    return pa.table([one(c) for c in columns], names=columns)


def supplier(sf: float, columns=None, strings: str = "view") -> pa.Table:
    n = counts(sf)["supplier"]
    a = {"s_suppkey": np.empty(n, "i8"), "s_nationkey": np.empty(n, "i8"), "s_acctbal": np.empty(n, "i8")}
    lib().tpch_gen_supplier(ctypes.c_double(sf), ctypes.c_int64(0), ctypes.c_int64(n),
                            _p(a["s_suppkey"]), _p(a["s_nationkey"]), _p(a["s_acctbal"]))
    columns = list(columns or list(a.keys()) + ["s_name"])

    def one(c):
        if c == "s_name":       # dbgen: "Supplier#%09d"
            txt = np.char.add("Supplier#", np.char.zfill(a["s_suppkey"].astype(str), 9))
            return pa.array(txt.tolist(), type=pa.string_view() if strings == "view" else pa.string())
        if c == "s_phone":
            return _phones(1, a["s_nationkey"], strings)
        if c == "s_address":
            return addresses(1, n, strings)
        if c == "s_comment":
            return comments(c, n, strings)
        return _to_arrow(c, a[c], strings)
    return pa.table([one(c) for c in columns], names=columns)


def partsupp(sf: float, columns=None, strings: str = "view") -> pa.Table:
    """ps_partkey, ps_suppkey, ps_availqty, ps_supplycost (ps_comment is not generated)"""
    n = counts(sf)["part"]
    a = {"ps_partkey": np.empty(4 * n, "i8"), "ps_suppkey": np.empty(4 * n, "i8"), "ps_availqty": np.empty(4 * n, "i4"), "ps_supplycost": np.empty(4 * n, "i8")}
    lib().tpch_gen_partsupp(ctypes.c_double(sf), ctypes.c_int64(0), ctypes.c_int64(n), _p(a["ps_partkey"]), _p(a["ps_suppkey"]),
                            _p(a["ps_availqty"]), _p(a["ps_supplycost"]))
    columns = list(columns or a.keys())
    return pa.table([decimal_from_int64(a[c]) if c == "ps_supplycost" else _to_arrow(c, a[c], strings) for c in columns], names=columns)


# dists.dss p_types / p_cntr: full strings in nested syllable order, equal weights
P_TYPES = [f"{a} {b} {c}" for a in ("STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO")
           for b in ("ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED") for c in ("TIN", "NICKEL", "BRASS", "STEEL", "COPPER")]
P_CONTAINERS = [f"{a} {b}" for a in ("SM", "LG", "MED", "JUMBO", "WRAP") for b in ("CASE", "BOX", "BAG", "JAR", "PACK", "PKG", "CAN", "DRUM")]   # PACK before PKG: part.tbl row 1 is "JUMBO PKG"


def part(sf: float, columns=None, strings: str = "view") -> pa.Table:
    """p_partkey, p_name, p_mfgr, p_brand ('Brand#MN'), p_type, p_size, p_container, p_retailprice (p_comment is not generated)"""
    n = counts(sf)["part"]
    a = {"p_partkey": np.empty(n, "i8"), "p_retailprice": np.empty(n, "i8"), "p_size": np.empty(n, "i4"), "p_type": np.empty(n, "u1"),
         "p_brand": np.empty(n, "i4"), "p_container": np.empty(n, "u1")}
    lib().tpch_gen_part(ctypes.c_double(sf), ctypes.c_int64(0), ctypes.c_int64(n), _p(a["p_partkey"]), _p(a["p_retailprice"]),
                        _p(a["p_size"]), _p(a["p_type"]), _p(a["p_brand"]), _p(a["p_container"]))
    columns = list(columns or ["p_partkey", "p_brand", "p_type", "p_size", "p_container", "p_retailprice"])
    out = []
    for c in columns:
        if c == "p_type":
            out.append(strings_from_codes(a[c], P_TYPES, strings))
        elif c == "p_container":
            out.append(strings_from_codes(a[c], P_CONTAINERS, strings))
        elif c == "p_name":
            out.append(part_names(n, strings))
        elif c == "p_mfgr":       # mk_part: brand = mfgr * 10 + U(1, 5)
            out.append(strings_from_codes(a["p_brand"] // 10 - 1, [f"Manufacturer#{m}" for m in range(1, 6)], strings))
        elif c == "p_brand":
            brands = sorted(set(int(x) for x in a[c]))
            codes = np.searchsorted(np.array(brands), a[c])
            out.append(strings_from_codes(codes, [f"Brand#{b}" for b in brands], strings))
        else:
            out.append(_to_arrow(c, a[c], strings))
    return pa.table(out, names=columns)


def nation(strings: str = "view") -> pa.Table:
    keys = np.arange(25, dtype=np.int64)
    names = strings_from_codes(np.arange(25), [n for n, _ in NATIONS], strings)
    return pa.table([pa.array(keys), names, pa.array(np.array([r for _, r in NATIONS], dtype=np.int64))],
                    names=["n_nationkey", "n_name", "n_regionkey"])


def region(strings: str = "view") -> pa.Table:
    return pa.table([pa.array(np.arange(5, dtype=np.int64)), strings_from_codes(np.arange(5), REGIONS, strings)],
                    names=["r_regionkey", "r_name"])


def tables(sf: float, strings: str = "view") -> dict:
    """All generated tables (small scale factors only)."""
    return {"lineitem": lineitem(sf, strings=strings), "orders": orders(sf, ORDERS_ALL + ["o_comment"], strings=strings),
            "customer": customer(sf, ["c_custkey", "c_nationkey", "c_acctbal", "c_mktsegment", "c_name", "c_phone", "c_address", "c_comment"], strings=strings),
            "supplier": supplier(sf, ["s_suppkey", "s_nationkey", "s_acctbal", "s_name", "s_phone", "s_address", "s_comment"], strings=strings),
            "part": part(sf, ["p_partkey", "p_brand", "p_type", "p_size", "p_container", "p_retailprice", "p_name", "p_mfgr"], strings=strings),
            "partsupp": partsupp(sf, strings=strings), "nation": nation(strings), "region": region(strings)}
