#!/bin/bash
# full ncu capture of the 14 forward launches of the second subject (MODE default fp16x2); report -> gpurun_out/full_$MODE.ncu-rep
MODE=${MODE:-fp16x2}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-'conv_|side_|head_'} -s ${KSKIP:-14} -c ${KCOUNT:-14} -f -o gpurun_out/full_$MODE python experiments/profile_one.py $MODE 2 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log; ls -la gpurun_out/full_$MODE.ncu-rep
