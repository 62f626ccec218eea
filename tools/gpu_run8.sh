#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
rm -f gpurun_out/ab2.log
for env in "X=1" "RRV_NO_TAIL_SPLIT=1" "X=2" "RRV_NO_TAIL_SPLIT=1"; do
  echo "== $env" >> gpurun_out/ab2.log
  env $env timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-side --no-bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'conv_ms',d['roofline']['kernel_ms_per_frame']); print([ (l['layer'][8:],l['ms']) for l in d['layers'] if '256' in l['layer'] or '512' in l['layer']])" >> gpurun_out/ab2.log 2>&1
done
cat gpurun_out/ab2.log
