# tile-shape / register-budget sweep for single operators (diagnostics): prints one line per (config, operator)
for cfg in "" "SAILGPU_RPT=2 SAILGPU_MINB=3" "SAILGPU_RPT=2 SAILGPU_MINB=4" "SAILGPU_RPT=1 SAILGPU_MINB=4" "SAILGPU_RPT=4 SAILGPU_STAGES=2"; do
  for op in join q6 high "q1 (sel"; do
    echo "== [$cfg] $op"
    env $cfg OPS_ONLY="$op" OPS_REPS=2 timeout 200 python scripts/bench_ops.py 10 2>&1 | grep -v "^{" | grep "Exec" | sed -e 's/"rows_in.*"GBps"/"GBps"/' | cut -c1-200
  done
done
