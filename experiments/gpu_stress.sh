#!/bin/bash
# stability check: the default bench N times under a short timeout (a hang shows as rc=124), then the c1 / c2 workloads
mkdir -p gpurun_out
for i in $(seq 1 ${REPS:-4}); do
  timeout ${BT:-150} python bench.py --subjects ${SUBJ:-32} --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/stress_$i.json 2> gpurun_out/stress_$i.err
  echo "run $i rc=$? $(cut -c1-120 gpurun_out/stress_$i.json)"
done
for wl in c1 c2; do
  timeout ${BT:-150} python bench.py --workload $wl --steps 3 --warmup 3 --cpu-frames 0 > gpurun_out/stress_$wl.json 2> gpurun_out/stress_$wl.err
  echo "$wl rc=$? $(cut -c1-120 gpurun_out/stress_$wl.json)"
done
