#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/layer_bench.py --one --kf kfup_32 kfup_32_3 2>&1 | tail -2
for r in 1 2 3; do
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/fr2_$r.json | python tools/benchline.py $([ $r = 1 ] && echo --layers)
done
