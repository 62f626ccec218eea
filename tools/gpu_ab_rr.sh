#!/bin/bash
# A/B of the staged row-reuse epilogue (RRV_NO_RSTAGE=1 = before): GPU tests, single layers, whole-frame bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_rr.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_rr.log
for v in 1 0; do
  if [ $v = 1 ]; then export RRV_NO_RSTAGE=1; else unset RRV_NO_RSTAGE; fi
  echo "== NO_RSTAGE=$v"
  python tools/layer_bench.py --one --kf kfup kfup3 2>&1 | tail -2
  python tools/layer_bench.py --one c64_128 c64_512 2>&1 | tail -2
done
for r in 1 2; do
for v in 1 0; do
  if [ $v = 1 ]; then export RRV_NO_RSTAGE=1; else unset RRV_NO_RSTAGE; fi
  echo "== frame NO_RSTAGE=$v"
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_rr_${v}_$r.json | python tools/benchline.py $([ $r = 1 ] && echo --layers)
done
done
