#!/bin/bash
# compute-sanitizer memcheck + racecheck (+ synccheck) on smoke() and two GPU tests of the default (fp16x2) mode; logs -> gpurun_out/
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # tool, name, command...
  tool=$1; name=$2; shift 2
  timeout ${ST:-900} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 "$@" > gpurun_out/sanitize_${tool}_${name}.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_${tool}_${name}.log
  echo "== $tool $name"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|passed|failed|smoke\[" gpurun_out/sanitize_${tool}_${name}.log | tail -6
}
run memcheck smoke python __graft_entry__.py smoke
run memcheck tests python -m pytest tests/test_gpu_tc.py -q -m gpu -k "test_forward_tc_small and fp16x2 or test_forward_x3_shapes and fp16x2 and 130" --tb=short
run racecheck smoke python __graft_entry__.py smoke
run racecheck tests python -m pytest tests/test_gpu_tc.py -q -m gpu -k "test_forward_tc_small and fp16x2 and 4" --tb=short
run synccheck smoke python __graft_entry__.py smoke
