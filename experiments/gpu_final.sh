#!/bin/bash
# Round-end GPU visit: parity tests, default bench (both arms), launch list, ncu --set full of one launch of every kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
  tail -4 gpurun_out/tests.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
SECONDS=0; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench.py default run: ${SECONDS}s wall" | tee gpurun_out/bench_wall.txt; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv --log-file gpurun_out/launches.csv python experiments/profile_one.py bf16 2 > gpurun_out/prof1.log 2>&1
python experiments/launch_summary.py gpurun_out/launches.csv 27 | tee gpurun_out/launch_summary.txt
if [ "${FULL:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 27 -f -o gpurun_out/full python experiments/profile_one.py bf16 2 > gpurun_out/prof2.log 2>&1
  ls -la gpurun_out/full.ncu-rep
fi
