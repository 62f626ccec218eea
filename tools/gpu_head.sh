#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 1 0; do
  if [ $v = 1 ]; then export RRV_NO_MERGE_WLO=1; else unset RRV_NO_MERGE_WLO; fi
  echo "== NO_MERGE_WLO=$v"; python tools/layer_bench.py --one head 2>&1 | tail -1
done
for r in 1 2; do for v in 1 0; do
  if [ $v = 1 ]; then export RRV_NO_MERGE_WLO=1; else unset RRV_NO_MERGE_WLO; fi
  echo "== frame NO_MERGE_WLO=$v"
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/head_${v}_$r.json | python tools/benchline.py --layers | grep -E "value|64->3 "
done; done
