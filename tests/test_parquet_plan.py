"""CPU test of the host half of the Parquet decoder (sail_b200/csrc/parquet.cu): the Thrift page-header reader and the walk over
RLE / bit-packed run headers must account for every value pyarrow's own metadata reports -- pages V1 and V2, dictionary and
plain pages, the writer's dictionary fallback, nulls."""
import io

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from sail_b200 import engine
from tests.test_gpu_parquet import table


@pytest.mark.parametrize("n", [1, 1000, 70001])
@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("version,use_dict,page", [("1.0", True, 1 << 20), ("1.0", False, 4096), ("2.0", True, 8192), ("2.0", False, 1 << 20)])
def test_page_and_run_walk_accounts_for_every_value(n, nulls, version, use_dict, page):
    t = table(n, 7 + n, nulls)
    buf = io.BytesIO()
    pq.write_table(t, buf, compression="none", use_dictionary=use_dict, data_page_version=version, data_page_size=page, dictionary_pagesize_limit=1 << 14)
    raw = buf.getvalue()
    for i, name in enumerate(t.schema.names):
        info = engine.parquet_inspect(raw, i)
        col = t.column(name)
        assert info["dense"] == n - col.null_count, (name, info)
        assert info["level_values"] == n, (name, info)            # every column is `optional`: one definition level per row
        if not use_dict:
            assert info["dict_pages"] == 0 and info["index_values"] == 0, (name, info)
        else:
            # dictionary pages carry indices for exactly their non-null values; after a fallback the rest is plain
            assert info["index_values"] <= info["dense"], (name, info)
            if not info["plain_pages"]:
                assert info["index_values"] == info["dense"] and info["dict_count"] == len(set(col.drop_null().to_pylist())), (name, info)
        if pa.types.is_string(col.type) and info["plain_pages"] and not info["dict_pages"]:
            assert info["plain_strings"] == info["dense"], (name, info)


def test_compressed_chunks_are_refused_at_plan_time():
    buf = io.BytesIO()
    pq.write_table(table(100, 1, False), buf, compression="snappy")
    with pytest.raises(engine.SailGpuError) as e:
        engine.parquet_inspect(buf.getvalue(), 0)
    assert e.value.code == 2
