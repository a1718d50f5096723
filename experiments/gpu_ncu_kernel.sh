#!/bin/bash
# full ncu capture of selected kernels (regex in $KREGEX), second subject
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-head_tc} -s ${KSKIP:-4} -c ${KCOUNT:-1} -f -o gpurun_out/kern python experiments/profile_one.py bf16 2 > gpurun_out/prof_kern.log 2>&1
tail -3 gpurun_out/prof_kern.log; ls -la gpurun_out/kern.ncu-rep
