#!/bin/bash
# One GPU-box visit: all GPU tests, smoke, default bench (c3) and the c1 / c2 / c4 workloads; outputs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
  tail -${TTAIL:-15} gpurun_out/tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
fi
for wl in ${WORKLOADS:-c3 c1 c2 c4}; do
  extra=""; [ "$wl" = "c3" ] && extra="--subjects ${SUBJ:-64}"
  timeout 900 python bench.py --workload $wl $extra --steps ${STEPS:-2} --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -3 gpurun_out/bench_$wl.err; cat gpurun_out/bench_$wl.json
done
