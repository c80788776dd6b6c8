"""Parquet column chunks decoded on the device (sail_b200/csrc/parquet.cu) against pyarrow's reader: data pages V1 and V2,
dictionary and plain encodings (and the writer's dictionary fallback), nulls, decimals as FIXED_LEN_BYTE_ARRAY and as
integers, short and long strings, several pages per chunk."""
import decimal
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

pytestmark = pytest.mark.gpu


def table(n, seed, nulls):
    rng = np.random.default_rng(seed)
    m = (lambda p: rng.random(n) < p) if nulls else (lambda p: None)
    words = ["alpha", "beta", "", "a considerably longer string value", "x" * 13, "twelve bytes"]
    return pa.table({
        "k": pa.array(rng.integers(0, 1 << 40, n), type=pa.int64(), mask=m(0.1)),
        "small": pa.array(rng.integers(0, 50, n).astype(np.int32), mask=m(0.05)),
        "d": pa.array([decimal.Decimal(int(x)) / 100 for x in rng.integers(-10**9, 10**9, n)], type=pa.decimal128(15, 2), mask=m(0.2)),
        "wide": pa.array([decimal.Decimal(int(x)) * 10**15 for x in rng.integers(-10**9, 10**9, n)], type=pa.decimal128(30, 0)),
        "s": pa.array([words[i] for i in rng.integers(0, len(words), n)], type=pa.string(), mask=m(0.1)),
        "u": pa.array([f"unique-{i:09d}-{'y' * (i % 7)}" for i in rng.permutation(n)], type=pa.string()),
        "dt": pa.array(rng.integers(8000, 11000, n).astype(np.int32), type=pa.int32()).cast(pa.date32()),
        "f": pa.array(rng.normal(size=n), mask=m(0.3)),
    })


def host(dev):
    from sail_b200 import engine
    spec = {"op": "projection", "exprs": [{"expr": {"col": i}, "name": n} for i, n in enumerate(dev.schema.names)]}
    op = engine.GpuExec(spec, [dev.schema])
    op.push(dev)
    op.finish()
    t = op.collect()
    op.close()
    return t


@pytest.mark.parametrize("n", [0, 1, 1000, 70001])
@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("version,use_dict,page", [("1.0", True, 1 << 20), ("1.0", False, 4096), ("2.0", True, 8192), ("2.0", False, 1 << 20)])
def test_parquet_decode_matches_pyarrow(n, nulls, version, use_dict, page):
    from sail_b200 import engine
    t = table(n, 7 + n, nulls)
    buf = io.BytesIO()
    pq.write_table(t, buf, compression="none", use_dictionary=use_dict, data_page_version=version, data_page_size=page, dictionary_pagesize_limit=1 << 14)
    raw = buf.getvalue()
    want = pq.read_table(io.BytesIO(raw))
    got = host(engine.parquet_decode(raw))
    assert got.num_rows == want.num_rows
    for name in want.schema.names:
        w = want.column(name).combine_chunks()
        g = got.column(name).combine_chunks()
        if pa.types.is_string(w.type):
            g = g.cast(pa.string())
        assert g.equals(w), name


def test_parquet_decimals_stored_as_integers_and_projection():
    from sail_b200 import engine
    t = table(5000, 3, True).select(["d", "k", "s"])
    buf = io.BytesIO()
    pq.write_table(t, buf, compression="none", store_decimal_as_integer=True)
    raw = buf.getvalue()
    got = host(engine.parquet_decode(raw, columns=["s", "d"]))
    assert got.schema.names == ["s", "d"]
    assert got.column("d").combine_chunks().equals(t.column("d").combine_chunks())
    assert got.column("s").combine_chunks().cast(pa.string()).equals(t.column("s").combine_chunks())


def test_parquet_compressed_pages_are_refused_loudly():
    from sail_b200 import engine
    buf = io.BytesIO()
    pq.write_table(table(100, 1, False), buf, compression="snappy")
    with pytest.raises(engine.SailGpuError) as e:
        engine.parquet_decode(buf.getvalue())
    assert e.value.code == 2
