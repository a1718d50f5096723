#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, one full ncu capture of a whole sub-batch.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
  tail -5 gpurun_out/tests.log
fi
timeout 600 python bench.py --subjects ${SUBJ:-16} --steps 2 --warmup 3 --cpu-frames ${CPUF:-0} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 97 -c 110 --csv --log-file gpurun_out/launches.csv python experiments/profile_one.py bf16 2 > gpurun_out/prof1.log 2>&1
python experiments/launch_summary.py gpurun_out/launches.csv 97 | tee gpurun_out/launch_summary.txt
if [ "${FULL:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -s ${FULL_SKIP:-106} -c ${FULL_COUNT:-22} -f -o gpurun_out/full python experiments/profile_one.py bf16 2 > gpurun_out/prof2.log 2>&1
  ls -la gpurun_out/
fi
