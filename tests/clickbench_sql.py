"""TEST INFRASTRUCTURE.  The ClickBench queries written a second time, straight from their SQL text
(python/pysail/data/clickbench/queries.sql) in pandas -- no operator specs, no oracle -- so that the plans of
sail_b200/clickbench.py (and the numpy oracle that executes them) are checked against an independent statement of what each
query means.  Every function returns the FULL result (no LIMIT / OFFSET), ordered by the query's ORDER BY; ties are in
arbitrary order, `tests.util.assert_topk` compares up to ties."""
import datetime

import numpy as np
import pandas as pd
import pyarrow as pa

JULY1, JULY31 = datetime.date(2013, 7, 1), datetime.date(2013, 7, 31)


def frame(hits: pa.Table) -> pd.DataFrame:
    cols = {}
    for name in hits.schema.names:
        c = hits.column(name)
        if pa.types.is_string_view(c.type):
            c = c.cast(pa.string())
        cols[name] = c.to_pandas()
    return pd.DataFrame(cols)


def counted(df, keys, name="c"):
    return df.groupby(keys, sort=False).size().reset_index(name=name)


def distinct_users(df, keys, name="u"):
    return df.groupby(keys, sort=False)["UserID"].nunique().reset_index(name=name)


def desc(df, key):
    return df.sort_values(key, ascending=False, kind="stable").reset_index(drop=True)


def one(**kw):
    return pd.DataFrame({k: [v] for k, v in kw.items()})


def july(df, counter=62):
    return df[(df.CounterID == counter) & (df.EventDate >= JULY1) & (df.EventDate <= JULY31)]


def q0(h): return one(**{"count(*)": len(h)})
def q1(h): return one(**{"count(*)": int((h.AdvEngineID != 0).sum())})
def q2(h): return one(**{"sum(AdvEngineID)": int(h.AdvEngineID.astype(np.int64).sum()), "count(*)": len(h), "avg(ResolutionWidth)": float(h.ResolutionWidth.astype(np.float64).mean())})
def q3(h): return one(**{"avg(UserID)": float(h.UserID.astype(np.float64).mean())})
def q4(h): return one(**{"count(DISTINCT UserID)": int(h.UserID.nunique())})
def q5(h): return one(**{"count(DISTINCT SearchPhrase)": int(h.SearchPhrase.nunique())})
def q6(h): return one(**{"min(EventDate)": h.EventDate.min(), "max(EventDate)": h.EventDate.max()})
def q7(h): return desc(counted(h[h.AdvEngineID != 0], ["AdvEngineID"], "count(*)"), "count(*)")
def q8(h): return desc(distinct_users(h, ["RegionID"]), "u")


def q9(h):
    g = h.groupby("RegionID", sort=False)
    out = pd.DataFrame({"sum(AdvEngineID)": g.AdvEngineID.sum().astype(np.int64), "c": g.size(), "avg(ResolutionWidth)": g.ResolutionWidth.mean().astype(np.float64),
                        "count(DISTINCT UserID)": g.UserID.nunique()}).reset_index()
    return desc(out, "c")


def q10(h): return desc(distinct_users(h[h.MobilePhoneModel != ""], ["MobilePhoneModel"]), "u")
def q11(h): return desc(distinct_users(h[h.MobilePhoneModel != ""], ["MobilePhone", "MobilePhoneModel"]), "u")
def q12(h): return desc(counted(h[h.SearchPhrase != ""], ["SearchPhrase"]), "c")
def q13(h): return desc(distinct_users(h[h.SearchPhrase != ""], ["SearchPhrase"]), "u")
def q14(h): return desc(counted(h[h.SearchPhrase != ""], ["SearchEngineID", "SearchPhrase"]), "c")
def q15(h): return desc(counted(h, ["UserID"], "count(*)"), "count(*)")
def q16(h): return desc(counted(h, ["UserID", "SearchPhrase"], "count(*)"), "count(*)")
def q17(h): return counted(h, ["UserID", "SearchPhrase"], "count(*)").sort_values(["UserID", "SearchPhrase"], kind="stable").reset_index(drop=True)
def q19(h, user): return h[h.UserID == user][["UserID"]].reset_index(drop=True)
def q20(h): return one(**{"count(*)": int(h.URL.str.contains("google", regex=False).sum())})
def q23(h, columns): return h[h.URL.str.contains("google", regex=False)][columns].sort_values("EventTime", kind="stable").reset_index(drop=True)
def q24(h): return h[h.SearchPhrase != ""][["EventTime", "SearchPhrase"]].sort_values("EventTime", kind="stable").reset_index(drop=True)
def q25(h): return h[h.SearchPhrase != ""][["SearchPhrase"]].sort_values("SearchPhrase", kind="stable", key=lambda s: s.str.encode("utf-8")).reset_index(drop=True)
def q26(h): return h[h.SearchPhrase != ""][["EventTime", "SearchPhrase"]].sort_values(["EventTime", "SearchPhrase"], kind="stable", key=lambda s: s.str.encode("utf-8") if s.dtype == object else s).reset_index(drop=True)


def q29(h, n=90):
    w = h.ResolutionWidth.astype(np.int64)
    return one(**{("sum(ResolutionWidth)" if i == 0 else f"sum((ResolutionWidth + {i}))"): int((w + i).sum()) for i in range(n)})


def q30_32(h, keys, filtered):
    if filtered:
        h = h[h.SearchPhrase != ""]
    g = h.groupby(keys, sort=False)
    out = pd.DataFrame({"c": g.size(), "sum(IsRefresh)": g.IsRefresh.sum().astype(np.int64), "avg(ResolutionWidth)": g.ResolutionWidth.mean().astype(np.float64)}).reset_index()
    return desc(out, "c")


def q30(h): return q30_32(h, ["SearchEngineID", "ClientIP"], True)
def q31(h): return q30_32(h, ["WatchID", "ClientIP"], True)
def q32(h): return q30_32(h, ["WatchID", "ClientIP"], False)
def q33(h): return desc(counted(h, ["URL"]), "c")


def q34(h):
    out = counted(h, ["URL"])
    out.insert(0, "1", np.int32(1))
    return desc(out, "c")


def q35(h):
    out = counted(h, ["ClientIP"])
    for i in (1, 2, 3):
        out.insert(i, f"(ClientIP - {i})", (out.ClientIP - i).astype(np.int32))
    return desc(out, "c")


def q36(h):
    f = july(h)
    return desc(counted(f[(f.DontCountHits == 0) & (f.IsRefresh == 0) & (f.URL != "")], ["URL"], "PageViews"), "PageViews")


def q37(h):
    f = july(h)
    return desc(counted(f[(f.DontCountHits == 0) & (f.IsRefresh == 0) & (f.Title != "")], ["Title"], "PageViews"), "PageViews")


def q38(h):
    f = july(h)
    return desc(counted(f[(f.IsRefresh == 0) & (f.IsLink != 0) & (f.IsDownload == 0)], ["URL"], "PageViews"), "PageViews")


def q39(h):
    f = july(h)
    f = f[f.IsRefresh == 0].copy()
    f["Src"] = np.where((f.SearchEngineID == 0) & (f.AdvEngineID == 0), f.Referer, "")
    f["Dst"] = f.URL
    return desc(counted(f, ["TraficSourceID", "SearchEngineID", "AdvEngineID", "Src", "Dst"], "PageViews"), "PageViews")


def q40(h, referer_hash):
    f = july(h)
    f = f[(f.IsRefresh == 0) & f.TraficSourceID.isin([-1, 6]) & (f.RefererHash == referer_hash)]
    return desc(counted(f, ["URLHash", "EventDate"], "PageViews"), "PageViews")


def q41(h, url_hash):
    f = july(h)
    f = f[(f.IsRefresh == 0) & (f.DontCountHits == 0) & (f.URLHash == url_hash)]
    return desc(counted(f, ["WindowClientWidth", "WindowClientHeight"], "PageViews"), "PageViews")
