#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export RRV_RSTAGE_NORES=1; else unset RRV_RSTAGE_NORES; fi
  echo "== RSTAGE_NORES=$v"
  python tools/layer_bench.py --one c64_128 c64_512 c128_128 2>&1 | tail -3
done
for r in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export RRV_RSTAGE_NORES=1; else unset RRV_RSTAGE_NORES; fi
  echo "== frame RSTAGE_NORES=$v"
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_rr2_${v}_$r.json | python tools/benchline.py
done
done
unset RRV_RSTAGE_NORES
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 9 -c 1 -f -o gpurun_out/r2b_kfup_rr python tools/profile_frame.py > gpurun_out/prof_kfup_rr.log 2>&1; echo "ncu rc=$?"
