#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
for l in 1 2; do
  echo "lanes=$l pdl"; python bench.py --steps 20 --warmup 5 --lanes $l --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
  echo "lanes=$l nopdl"; RRV_NO_PDL=1 python bench.py --steps 20 --warmup 5 --lanes $l --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
done
done
echo "lanes=2 nopdl 200"; RRV_NO_PDL=1 python bench.py --steps 200 --warmup 10 --lanes 2 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py | cut -c1-60
