#!/bin/bash
# A/B bench of engine switches (environment variables read at engine creation): each line of $AB is one configuration
mkdir -p gpurun_out
: > gpurun_out/ab.log
while IFS= read -r cfg; do
  [ -z "$cfg" ] && continue
  echo "== $cfg" >> gpurun_out/ab.log
  env $cfg timeout 600 python bench.py --subjects ${SUBJ:-32} --steps 2 --warmup 3 --cpu-frames 0 2>>gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f fwd_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline'].get('whole_forward', d['roofline'])['avg_forward_ms_per_subject']))" >> gpurun_out/ab.log
done <<< "$AB"
cat gpurun_out/ab.log
