#!/bin/bash
# compute-sanitizer passes over a subset of the GPU tests (the subset keeps the run under a few minutes); output in gpurun_out/
# usage (GPU box): bash profiles/tools/sanitize.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
SEL='swept_end_to_end_mesh and lprism or batched_device_callback or swept_golden or discrete_golden or discrete_parity_mesh and box or shards_sum_to_full and 3 or frontend_kernels and TwistBox or frontend_state or frontend_two_pass or device_lockstep_lbfgs or batched_callback_with_swept or open_mesh_discrete_cost or config0_ball'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 \
      python -m pytest tests -m gpu -q -x --timeout 800 -k "$SEL" > gpurun_out/sanitize_${tag}_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tag}_${tool}.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tag}_${tool}.log | sort | uniq -c | tr '\n' ';')  $(tail -1 gpurun_out/sanitize_${tag}_${tool}.log)"
done
