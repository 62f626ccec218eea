#!/bin/bash
# A/B of the epilogue residual prefetch variants (tools/build_variants.sh): single layers, then whole-frame bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CS=rerevst-code_b200/csrc
for v in base l2pf1 l2pf3 pre1; do
  echo "== $v"
  RRV_LIB_PATH=$CS/librerevst_b200_$v.so python tools/layer_bench.py --one --kf kfup kfup3 2>&1 | tail -2
  RRV_LIB_PATH=$CS/librerevst_b200_$v.so python tools/layer_bench.py --one --full c64_64 c256_256 c128_128 2>&1 | tail -3
done
for r in 1 2; do
for v in base l2pf1 l2pf3 pre1; do
  echo "== frame $v"
  RRV_LIB_PATH=$CS/librerevst_b200_$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_${v}_$r.json | python tools/benchline.py
done
done
