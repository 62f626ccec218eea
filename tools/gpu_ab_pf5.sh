#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CS=rerevst-code_b200/csrc
for r in 1 2; do
for v in "" _pf5; do
  echo "== frame lib$v"
  RRV_LIB_PATH=$CS/librerevst_b200$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | tee gpurun_out/ab_pf5${v}_$r.json | python tools/benchline.py --layers | grep -E "value|32->512"
done
done
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 9 -c 1 -f -o gpurun_out/r2b_kfup_sw64 python tools/profile_frame.py > gpurun_out/prof_kfup_sw64.log 2>&1; echo "ncu rc=$?"
