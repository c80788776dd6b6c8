"""oracle/render.py -- TEST INFRASTRUCTURE.  Renders a result table the way the reference's
snapshot tests do (python/pysail/testing/spark/utils/sql.py parse/format_show_string): every cell
as Spark's show() string -- decimals at full scale, dates ISO, NULL for nulls."""
import datetime
import decimal

import pyarrow as pa


def cell(v, t) -> str:
    if v is None:
        return "NULL"
    if pa.types.is_decimal(t):
        return format(v, "f") if isinstance(v, decimal.Decimal) else str(v)
    if pa.types.is_date32(t):
        return v.isoformat() if isinstance(v, datetime.date) else str(v)
    if pa.types.is_boolean(t):
        return "true" if v else "false"
    return str(v)


def rows(tbl: pa.Table):
    cols = [tbl.column(i).to_pylist() for i in range(tbl.num_columns)]
    types = [f.type for f in tbl.schema]
    return [[cell(c[r], t) for c, t in zip(cols, types)] for r in range(tbl.num_rows)]
