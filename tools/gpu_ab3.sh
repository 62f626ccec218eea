#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/ab3.log
for env in "X=1" "RRV_NO_TAIL_SPLIT=1" "X=2" "RRV_NO_TAIL_SPLIT=1" "X=3" "RRV_NO_TAIL_SPLIT=1"; do
  echo "== $env" >> gpurun_out/ab3.log
  env $env timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side --no-bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'conv_ms',round(d['roofline']['kernel_ms_per_frame'],4)); print([ (l['layer'][8:],l['ms']) for l in d['layers'] if '256->256' in l['layer'] or '256->512' in l['layer']])" >> gpurun_out/ab3.log 2>&1
  sleep 3
done
cat gpurun_out/ab3.log
