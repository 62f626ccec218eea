#!/bin/bash
# Full GPU test suite + two 20-step bench lines (the driver's own command shape) with the per-layer table of the first.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_line1.json | python tools/benchline.py --layers
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_line2.json | python tools/benchline.py
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_line3.json | python tools/benchline.py
