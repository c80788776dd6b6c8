# compute-sanitizer over a slice of the GPU parity suite (interpreter AND specialised kernels): memcheck for out-of-bounds /
# misaligned accesses, racecheck for shared-memory hazards in the tile pipeline (TMA stages, dictionary, look-back) and synccheck.
# Logs land in gpurun_out/evidence/; the passing ones are copied to profiles/.
set -x
OUT=gpurun_out/evidence; mkdir -p $OUT
R=${ROUND:-r02}
SEL='test_projection and 1000 or test_filter and 5000 and not two_pass or test_aggregate_single and 1000 or test_two_phase_aggregate or test_divide_by_zero or test_aggregate_batches_differ and alternating'
REL='test_hash_join_unique_build and 50-500 or test_hash_join_duplicate_build_keys or test_hash_join_string_keys or test_sort and 5000 or test_hash_repartition and 2 or test_row_round_robin_reference_kat or test_substr or test_nested_loop or test_sort_preserving_merge and 2-None or test_topk_selection and 10-keys0 or test_semi_anti_join_with_residual_filter and False'
for tool in memcheck racecheck synccheck; do
  SAILGPU_JIT_MIN_ROWS=0 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 0 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "$SEL" > $OUT/${R}_sanitizer_${tool}_pipeline_jit.log 2>&1; echo "exit=$?" >> $OUT/${R}_sanitizer_${tool}_pipeline_jit.log
  SAILGPU_JIT=0 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 0 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "$SEL" > $OUT/${R}_sanitizer_${tool}_pipeline_vm.log 2>&1; echo "exit=$?" >> $OUT/${R}_sanitizer_${tool}_pipeline_vm.log
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 0 python -m pytest tests/test_gpu_relational.py -m gpu -x -q -k "$REL" > $OUT/${R}_sanitizer_${tool}_relational.log 2>&1; echo "exit=$?" >> $OUT/${R}_sanitizer_${tool}_relational.log
done
tail -n 6 $OUT/${R}_sanitizer_*.log
