#!/bin/bash
# quick GPU visit: selected tests + launch list + short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${TESTSEL:-} > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
tail -15 gpurun_out/tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-97} -c 110 --csv --log-file gpurun_out/launches.csv python experiments/profile_one.py bf16 2 > gpurun_out/prof1.log 2>&1
python experiments/launch_summary.py gpurun_out/launches.csv ${LN:-97} | tee gpurun_out/launch_summary.txt
timeout 600 python bench.py --subjects ${SUBJ:-16} --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
