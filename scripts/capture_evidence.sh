# Round evidence on one B200: the bench line (not under a profiler), the reference arm, the ncu launch list of the same
# bench command, one `ncu --set full` capture of the dominant kernel (fused Q1 pipeline) as CSV pages, operator table.
# Everything lands in gpurun_out/evidence/ (small files); the curated copies live in profiles/.
set -x
OUT=gpurun_out/evidence; mkdir -p $OUT
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_q1_bench.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > $OUT/ncu_launches.log 2>&1
# launches of pipeline_kernel in run_q1_once.py: [0] table import, then per repetition {fused Q1, final aggregate}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pipeline_kernel --launch-skip 3 -c 1 -o /tmp/q1 python scripts/run_q1_once.py 10 3 > $OUT/ncu_q1.log 2>&1
ncu -i /tmp/q1.ncu-rep --page raw --csv > $OUT/pipeline_kernel_q1_raw.csv 2>/dev/null
ncu -i /tmp/q1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/q1_source.csv 2>/dev/null
python scripts/ncu_source_summary.py /tmp/q1_source.csv 0 60 > $OUT/pipeline_kernel_q1_source_top.txt
timeout 300 python scripts/bench_ops.py 10 > $OUT/ops.log 2>/dev/null; tail -1 $OUT/ops.log > $OUT/ops_sf10.json
timeout 600 python scripts/bench_tpch.py 10 > $OUT/tpch_sf10.txt 2>&1
ls -la $OUT
