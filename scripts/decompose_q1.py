"""Times variants of the Q1 pipeline to attribute kernel time to: tile loading, VM, key lookup, accumulation."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SAILGPU_TIMING"] = "1"
import bench  # noqa: E402
from sail_b200 import engine, plans  # noqa: E402


def time_spec(ctx, spec, dev, schema, reps=4):
    tot = n = 0
    for i in range(reps + 1):
        op = engine.GpuExec(spec, [schema], ctx)
        op.push(dev.borrow())
        op.finish()
        op.collect_device()
        m = op.metrics()
        op.close()
        if i:
            tot += m["gpu.pipeline_kernel_ns"]
            n += m["gpu.pipeline_launches"]
    return tot / 1e6 / max(1, n)


def main():
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    ctx = engine.Context(0)
    table = bench.gen_shard(sf, 0, 1).combine_chunks()
    dev = engine.to_device(table, ctx)
    names = table.schema.names
    fused, _, _ = bench.q1_specs()
    flt, proj, agg = fused["stages"]
    C = lambda n: {"col": names.index(n)}  # noqa: E731
    def aggspec(keys, aggs, schema_names):
        return {"op": "aggregate", "mode": "partial", "group_by": [{"expr": {"col": schema_names.index(k)}, "name": k} for k in keys], "aggs": aggs}
    variants = {}
    cnt = [{"fn": "count", "args": [], "name": "c"}]
    variants["V1 filter(shipdate)+count(*) nokeys [4 B/row]"] = {"op": "pipeline", "stages": [flt, aggspec([], cnt, ["q", "p", "d", "t", "rf", "ls"])]}
    variants["V2 filter+sum(qty) nokeys [20 B/row]"] = {"op": "pipeline", "stages": [flt, aggspec([], [{"fn": "sum", "args": [{"col": 0}], "name": "s"}], [])]}
    a_nokeys = copy.deepcopy(agg); a_nokeys["group_by"] = []
    variants["V4 full Q1 exprs, NO group keys (1 group, 6 accs) [68 B/row]"] = {"op": "pipeline", "stages": [flt, proj, a_nokeys]}
    a_cnt = copy.deepcopy(agg); a_cnt["aggs"] = cnt
    variants["V5 filter + group keys + count(*) only [36 B/row]"] = {"op": "pipeline", "stages": [flt, proj, a_cnt]}
    a_one = copy.deepcopy(agg); a_one["aggs"] = agg["aggs"][:1]
    variants["V5b keys + sum(qty) [52 B/row]"] = {"op": "pipeline", "stages": [flt, proj, a_one]}
    variants["V6 full Q1 [100 B/row]"] = fused
    q6 = plans.q6().inputs[0].inputs[0]
    st, n = [], q6
    while n.spec["op"] != "scan":
        st.append(n.spec); n = n.inputs[0]
    # q6 plan is over 4 columns in its own order: remap by building it against the 7-column table
    for name, spec in variants.items():
        for cfg in (dict(SAILGPU_RPT="1", SAILGPU_STAGES="2"), dict(SAILGPU_RPT="2", SAILGPU_STAGES="1"), dict(SAILGPU_RPT="4", SAILGPU_STAGES="1")):
            for k in ("SAILGPU_RPT", "SAILGPU_STAGES"):
                os.environ.pop(k, None)
            os.environ.update(cfg)
            try:
                ms = time_spec(ctx, spec, dev, table.schema)
                print(f"{name:62s} rpt={cfg['SAILGPU_RPT']} st={cfg['SAILGPU_STAGES']}: {ms:7.3f} ms", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name}: {cfg} FAILED {e}", flush=True)
    del dev
    ctx.synchronize()


if __name__ == "__main__":
    main()
