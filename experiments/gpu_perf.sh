#!/bin/bash
# perf visit: short bench + per-subject launch list (ncu gpu__time_duration) for MODE (default fp16x2)
MODE=${MODE:-fp16x2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 600 python bench.py --mode $MODE --subjects ${SUBJ:-16} --steps 2 --warmup 3 --cpu-frames ${CPUF:-0} > gpurun_out/bench_$MODE.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_$MODE.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$MODE.csv python experiments/profile_one.py $MODE 2 > gpurun_out/prof1.log 2>&1
python experiments/launch_last.py gpurun_out/launches_$MODE.csv | tee gpurun_out/launch_summary_$MODE.txt
