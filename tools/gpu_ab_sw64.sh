#!/bin/bash
# GPU tests on the 32-channel operand path, single-layer times (padded-to-64 vs 32-channel KernelFilter convolutions), frame lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_sw64.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_sw64.log
python tools/layer_bench.py --one --kf kfup kfup3 kfup_32 kfup_32_3 2>&1 | tail -4
python tools/layer_bench.py --one c512_64 kfdown_32 2>&1 | tail -2
for r in 1 2; do
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_sw64_$r.json | python tools/benchline.py $([ $r = 1 ] && echo --layers)
done
