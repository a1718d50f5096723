#!/bin/bash
# quick GPU iteration: tensor-core tests, one line per failure, output to gpurun_out/
mkdir -p gpurun_out
timeout ${T:-1500} python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=${TB:-line} ${K:+-k "$K"} ${X:+-x} 2>&1 | tail -${TAIL:-80} > gpurun_out/quick.log
cat gpurun_out/quick.log
