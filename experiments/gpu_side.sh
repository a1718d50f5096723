#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q -k "side_kernel or fused_and or sa_random" > gpurun_out/tests_side.log 2>&1; echo "rc=$?" >> gpurun_out/tests_side.log
tail -15 gpurun_out/tests_side.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-97} -c 110 --csv --log-file gpurun_out/launches.csv python experiments/profile_one.py bf16 2 > gpurun_out/prof1.log 2>&1
python experiments/launch_summary.py gpurun_out/launches.csv ${LN:-97} | tee gpurun_out/launch_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --subjects 16 --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
