"""The reference's small known-answer tests, restated as operator specs: the same inputs, the same expected rows.
Sources: python/pysail/tests/spark/test_group_by.py:10-73, test_sort.py:5-61, test_shuffle.py:8-46, test_join.txt:6-55,
test_repartition.py:57-81.  Each case = (name, spec-tree builder, expected rows as python tuples, ordered?)."""
import pyarrow as pa

from sail_b200 import plans as P

DEALER = pa.table({"id": pa.array([100, 100, 100, 200, 200, 200, 300, 300], type=pa.int32()),
                   "city": pa.array(["Fremont"] * 3 + ["Dublin"] * 3 + ["San Jose"] * 2),
                   "car_model": pa.array(["Honda Civic", "Honda Accord", "Honda CRV"] * 2 + ["Honda Civic", "Honda Accord"]),
                   "quantity": pa.array([10, 15, 7, 20, 10, 3, 5, 8], type=pa.int32())})
PERSON = pa.table({"id": pa.array([100, 200, 300, 400], type=pa.int32()), "name": pa.array(["Mary", "John", "Mike", "Dan"]),
                   "age": pa.array([None, 30, 80, 50], type=pa.int32())})
CUSTOMER = pa.table({"id": pa.array([1, 2, 3], type=pa.int64()), "name": pa.array(["Alice", "Bob", "Charlie"]), "age": pa.array([34, 36, 30], type=pa.int64())})
ORDER = pa.table({"oid": pa.array([1, 2, 3], type=pa.int64()), "customer_id": pa.array([1, 2, 2], type=pa.int64()), "amount": pa.array([100, 200, 300], type=pa.int64())})
TAB = pa.table({"a": pa.array([False, False, None], type=pa.bool_()), "b": pa.array([1, None, 3], type=pa.int32()), "c": pa.array([None, 2.0, 3.0], type=pa.float64())})
DF1 = pa.table({"name": pa.array(["Alice", "Bob"]), "age": pa.array([2, 5], type=pa.int64())})
DF2 = pa.table({"name2": pa.array(["Tom", "Bob"]), "height": pa.array([80, 85], type=pa.int64())})
DF3 = pa.table({"name3": pa.array(["Alice", "Bob", "Tom", None]), "age3": pa.array([10, 5, None, None], type=pa.int64()), "height3": pa.array([80, None, None, None], type=pa.int64())})

TABLES = {"dealer": DEALER, "person": PERSON, "customer": CUSTOMER, "order": ORDER, "tab": TAB, "df1": DF1, "df2": DF2, "df3": DF3}


def _sort(child, keys, fetch=None):
    """Spark default null ordering: ASC NULLS FIRST, DESC NULLS LAST"""
    return P.sort(child, keys, fetch)


def cases():
    dealer = P.scan("dealer", DEALER.schema.names)
    out = []
    # test_group_by.py:42-47  SELECT id, sum(quantity) FROM dealer GROUP BY id ORDER BY id
    g = P.two_phase(dealer, ["id"], [("sum", P.col("quantity"), "sum(quantity)", "Int32")])
    out.append(("group_by", _sort(g, [("id", True)]), [(100, 32), (200, 33), (300, 13)], True))
    # test_group_by.py:58-69  sum + max
    g = P.two_phase(dealer, ["id"], [("sum", P.col("quantity"), "sum", "Int32"), ("max", P.col("quantity"), "max", "Int32")])
    out.append(("multiple_aggregations", _sort(g, [("id", True)]), [(100, 32, 15), (200, 33, 20), (300, 13, 8)], True))
    # test_group_by.py:72-84  count(DISTINCT city) per car_model == group by (car_model, city) then count per car_model
    d = P.two_phase(dealer, ["car_model", "city"], [])
    g = P.two_phase(d, ["car_model"], [("count", None, "count", None)])
    out.append(("count_distinct", g, [("Honda Accord", 3), ("Honda CRV", 2), ("Honda Civic", 3)], False))
    # test_sort.py:5-61 over (false,1,NULL), (false,NULL,2.0), (NULL,3,3.0)
    tab = P.scan("tab", TAB.schema.names)
    out.append(("sort_a", P.project(_sort(tab, [("a", True)]), ["a"]), [(None,), (False,), (False,)], True))
    out.append(("sort_c", _sort(tab, [("c", True)]), [(False, 1, None), (False, None, 2.0), (None, 3, 3.0)], True))
    out.append(("sort_b", _sort(tab, [("b", True)]), [(False, None, 2.0), (False, 1, None), (None, 3, 3.0)], True))
    out.append(("sort_c_b", _sort(tab, [("c", True), ("b", True)]), [(False, 1, None), (False, None, 2.0), (None, 3, 3.0)], True))
    out.append(("sort_c_desc_b", _sort(tab, [("c", False), ("b", True)]), [(None, 3, 3.0), (False, None, 2.0), (False, 1, None)], True))
    out.append(("sort_c_desc_a", _sort(tab, [("c", False), ("a", True)]), [(None, 3, 3.0), (False, None, 2.0), (False, 1, None)], True))
    # test_shuffle.py:33-46  join + group by + order by total desc  => Bob 500, Alice 100
    j = P.hash_join(P.scan("customer", ["id", "name"]), P.scan("order", ["customer_id", "amount"]), [("id", "customer_id")], projection=["name", "amount"])
    g = P.two_phase(j, ["name"], [("sum", P.col("amount"), "total", "Int64")])
    out.append(("join_group_by", _sort(g, [("total", False)]), [("Bob", 500), ("Alice", 100)], True))
    # test_join.txt:16-21  df1.join(df2, "name") => Bob 5 85
    j = P.hash_join(P.scan("df1", ["name", "age"]), P.scan("df2", ["name2", "height"]), [("name", "name2")], projection=["name", "age", "height"])
    out.append(("join_using_name", j, [("Bob", 5, 85)], False))
    # test_join.txt:43-48  df1.join(df3, ["name", "age"]) => Bob 5 NULL   (two keys; NULL keys never match)
    j = P.hash_join(P.scan("df1", ["name", "age"]), P.scan("df3", ["name3", "age3", "height3"]), [("name", "name3"), ("age", "age3")],
                    projection=["name", "age", "height3"])
    out.append(("join_two_keys", j, [("Bob", 5, None)], False))
    # test_join.txt (left outer shape of the 'outer' doctest restricted to the build side): df2 right-outer from df1
    j = P.hash_join(P.scan("df1", ["name", "age"]), P.scan("df2", ["name2", "height"]), [("name", "name2")], join_type="right", projection=["name", "age", "name2", "height"])
    out.append(("join_right_outer", j, [("Bob", 5, "Bob", 85), (None, None, "Tom", 80)], False))
    return out


def run(run_op):
    """executes every case with `run_op(spec, *tables)`; returns [(name, got rows, expected rows, ordered)]"""
    res = []
    for name, plan, want, ordered in cases():
        got = P.execute(plan, TABLES, run_op)
        rows = [tuple(r.values()) for r in got.to_pylist()]
        res.append((name, rows, want, ordered))
    return res


def check(results):
    for name, got, want, ordered in results:
        g, w = (got, want) if ordered else (sorted(got, key=repr), sorted(want, key=repr))
        assert g == w, f"reference KAT '{name}': got {g}, expected {w}"
