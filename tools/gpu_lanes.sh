#!/bin/bash
# Frames in flight: tests of the lanes, then bench A/B (--lanes 1 / 2 / 3) at the driver's 20 steps and at the 200-step default.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "lanes or transfer_stream or facade or generate_real" 2>&1 | tail -3
for rep in 1 2; do
for l in 1 2 3; do
  echo "lanes=$l steps=20"; python bench.py --steps 20 --warmup 5 --lanes $l --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/lanes${l}_20.json | python tools/benchline.py | cut -c1-90
done
done
for l in 1 2 3; do
  echo "lanes=$l steps=200"; python bench.py --steps 200 --warmup 10 --lanes $l --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/lanes${l}_200.json | python tools/benchline.py | cut -c1-90
done
