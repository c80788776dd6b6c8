# profiles single operators of scripts/bench_ops.py under ncu and leaves CSV pages (raw + source) in gpurun_out/
# the first 2 pipeline_kernel launches of bench_ops.py are the host->device imports of its tables
NCU="ncu --set full --import-source on --clock-control none -k regex:pipeline_kernel --launch-skip 2"
prof() {  # name only count
  OPS_ONLY="$2" OPS_REPS=0 timeout 600 $NCU -c $3 -o /tmp/$1 python scripts/bench_ops.py 10 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
}
prof agg_high high 1

prof fq6 q6 1
ls -la gpurun_out/
