"""The result sink: sailgpu_ipc_stream frames a host batch as the self-contained Arrow IPC stream Sail's
Spark Connect executor sends per result batch (crates/sail-spark-connect/src/executor.rs:320-330 `to_arrow_batch`).  The metadata
flatbuffers are written by hand (sail_b200/csrc/ipc.cpp); pyarrow's stream reader -- which verifies them -- is the checker.
Host-only: runs without a GPU."""
import datetime
import decimal

import numpy as np
import pyarrow as pa
import pytest

from sail_b200 import engine


def roundtrip(batch: pa.RecordBatch) -> pa.Table:
    data = engine.ipc_stream(batch)
    reader = pa.ipc.open_stream(data)
    assert reader.schema == batch.schema
    batches = list(reader)
    assert len(batches) == 1                     # Schema, ONE RecordBatch, end of stream: what to_arrow_batch writes
    assert data[-8:] == b"\xff\xff\xff\xff\x00\x00\x00\x00"
    return pa.Table.from_batches(batches)


def sample(n, seed=0, nulls=True):
    rng = np.random.default_rng(seed)
    mask = (lambda: rng.random(n) < 0.2) if nulls else (lambda: None)
    words = ["", "a", "Customer#000000001", "special requests sleep furiously", "дом", "x" * 300]
    strs = [words[i] for i in rng.integers(0, len(words), n)]
    return pa.record_batch({
        "i8": pa.array(rng.integers(-128, 128, n).astype(np.int8), mask=mask()),
        "u8": pa.array(rng.integers(0, 256, n).astype(np.uint8), mask=mask()),
        "i16": pa.array(rng.integers(-2**15, 2**15, n).astype(np.int16), mask=mask()),
        "u16": pa.array(rng.integers(0, 2**16, n).astype(np.uint16), mask=mask()),
        "i32": pa.array(rng.integers(-2**31, 2**31, n).astype(np.int32), mask=mask()),
        "u32": pa.array(rng.integers(0, 2**32, n).astype(np.uint32), mask=mask()),
        "i64": pa.array(rng.integers(-2**62, 2**62, n).astype(np.int64), mask=mask()),
        "u64": pa.array(rng.integers(0, 2**63, n).astype(np.uint64), mask=mask()),
        "f32": pa.array(rng.random(n).astype(np.float32), mask=mask()),
        "f64": pa.array(rng.random(n), mask=mask()),
        "flag": pa.array(rng.random(n) < 0.5, mask=mask()),
        "money": pa.array([None if (nulls and rng.random() < 0.2) else decimal.Decimal(int(v)).scaleb(-2) for v in rng.integers(-10**14, 10**14, n)], pa.decimal128(15, 2)),
        "wide": pa.array([decimal.Decimal(int(v) * 10**19).scaleb(-4) for v in rng.integers(-10**14, 10**14, n)], pa.decimal128(38, 4)),
        "day": pa.array(rng.integers(0, 20000, n).astype(np.int32), pa.int32(), mask=mask()).cast(pa.date32()),
        "utf8": pa.array(strs, pa.string(), mask=mask()),
        "view": pa.array(strs, pa.string_view(), mask=mask()),
        "large": pa.array(strs, pa.large_string(), mask=mask()),
        "bin": pa.array([s.encode() for s in strs], pa.binary(), mask=mask()),
    })


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1000, 70001])
@pytest.mark.parametrize("nulls", [False, True])
def test_every_column_type_survives_the_round_trip(n, nulls):
    batch = sample(n, seed=n, nulls=nulls)
    assert roundtrip(batch).equals(pa.Table.from_batches([batch]))


def test_schema_only_stream():
    schema = sample(3).schema
    reader = pa.ipc.open_stream(engine.ipc_stream(None, schema))
    assert reader.schema == schema and list(reader) == []


@pytest.mark.parametrize("first,count", [(0, 10), (3, 50), (8, 64), (13, 1), (95, 5), (100, 0)])
def test_sliced_columns_are_rebased(first, count):
    """a consumer may hand over slices (non-zero offset): bitmaps are re-packed from the bit offset, Utf8 offsets rebased to zero"""
    whole = sample(100, seed=5)
    part = pa.record_batch([c.slice(first, count) for c in whole.columns], schema=whole.schema)
    assert roundtrip(part).equals(pa.Table.from_batches([part]))


def test_nullability_and_names_are_kept():
    schema = pa.schema([pa.field("k", pa.int64(), nullable=False), pa.field("имя", pa.string_view(), nullable=True), pa.field("", pa.date32())])
    batch = pa.record_batch([pa.array([1, 2]), pa.array(["a", None], pa.string_view()), pa.array([datetime.date(1995, 3, 15), None])], schema=schema)
    out = roundtrip(batch)
    assert [(f.name, f.nullable) for f in out.schema] == [("k", False), ("имя", True), ("", True)]
    assert out.equals(pa.Table.from_batches([batch]))


def test_views_with_several_data_buffers():
    a = pa.array(["first buffer holds this long string", "tiny"], pa.string_view())
    b = pa.array(["second buffer holds another long string", None], pa.string_view())
    col = pa.concat_arrays([a, b])
    assert len(col.buffers()) >= 4               # validity, views, two data buffers
    batch = pa.record_batch([col], names=["v"])
    assert roundtrip(batch).equals(pa.Table.from_batches([batch]))


def test_a_tpch_result_is_what_pyarrow_would_have_written(tpch_tiny):
    """same decoded content as pyarrow's own StreamWriter for a real result shape (Q1's ten columns at SF0.001, via the oracle)"""
    from sail_b200 import plans
    from tests.util import oracle_op
    res = plans.execute(plans.TPCH["q1"](), tpch_tiny, oracle_op).combine_chunks().to_batches()[0]
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, res.schema) as w:
        w.write_batch(res)
    theirs = pa.ipc.open_stream(sink.getvalue()).read_all()
    assert roundtrip(res).equals(theirs)


def test_unsupported_columns_are_refused():
    nested = pa.record_batch([pa.array([[1, 2], [3]])], names=["l"])
    with pytest.raises(engine.SailGpuError) as e:
        engine.ipc_stream(nested)
    assert e.value.code == 2
    ts = pa.record_batch([pa.array([1, 2], pa.timestamp("us"))], names=["t"])
    with pytest.raises(engine.SailGpuError) as e:
        engine.ipc_stream(ts)
    assert e.value.code == 2 and "tsu" in str(e.value)


@pytest.mark.parametrize("first,count", [(0, 1000), (3, 777), (64, 128), (5, 3)])
def test_unknown_null_count_is_counted(first, count):
    """a producer may leave null_count = -1 (Arrow C Data Interface): the RecordBatch's FieldNode needs the real count"""
    import ctypes
    rng = np.random.default_rng(first)
    col = pa.array(rng.integers(0, 100, 1200), mask=rng.random(1200) < 0.3).slice(first, count)
    batch = pa.record_batch([col], names=["x"])
    sc, c = engine._export_schema(batch.schema), engine.ArrowArrayC()
    batch._export_to_c(ctypes.addressof(c))
    child = engine.ArrowArrayC.from_address(ctypes.cast(c.children, ctypes.POINTER(ctypes.c_void_p))[0])
    child.null_count = -1
    data, n = ctypes.c_void_p(), ctypes.c_size_t(0)
    assert engine.lib().sailgpu_ipc_stream(ctypes.addressof(sc), ctypes.addressof(c), ctypes.byref(data), ctypes.byref(n)) == 0
    out = pa.ipc.open_stream(ctypes.string_at(data.value, n.value)).read_all()
    engine.lib().sailgpu_ipc_free(data)
    engine._release_schema(sc)
    ctypes.CFUNCTYPE(None, ctypes.c_void_p)(c.release)(ctypes.addressof(c))
    assert out.column(0).null_count == col.null_count and out.equals(pa.Table.from_batches([batch]))
