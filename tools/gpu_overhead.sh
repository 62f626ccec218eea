#!/bin/bash
# Per-frame fixed cost of the 25-launch graph: a 256 x 256 frame (padded 384 x 384: one or two tiles per worker per layer).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for ln in 1 2; do
  echo "== 256 lanes=$ln"; python bench.py --size 256 --lanes $ln --steps 200 --warmup 10 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py
done
echo "== 256 lanes=1 no PDL"; RRV_NO_PDL=1 python bench.py --size 256 --lanes 1 --steps 200 --warmup 10 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py
for ln in 1 2; do
  echo "== 1080p lanes=$ln"; python bench.py --lanes $ln --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py
done
