#!/bin/bash
# Bounding experiments for the epilogue of the merged-tap layers: build the measurement-only variants first
#   tools/build_variants.sh nostore "-DRRV_EXP_NOSTORE=1" notab "-DRRV_EXP_NOTAB=1" noshfl "-DRRV_EXP_NOSHFL=1" all3 "-DRRV_EXP_NOSTORE=1 -DRRV_EXP_NOTAB=1 -DRRV_EXP_NOSHFL=1"
# (their results are wrong by construction; profiles/r2_epilogue_bounding_experiments.txt)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
D=rerevst-code_b200/csrc
for v in "" _nostore _notab _noshfl _all3; do
  echo "== variant '$v'" >> gpurun_out/exp1.log
  RRV_LIB_PATH=$PWD/$D/librerevst_b200$v.so timeout 300 python tools/layer_bench.py --small --full c64_64 >> gpurun_out/exp1.log 2>&1
  RRV_LIB_PATH=$PWD/$D/librerevst_b200$v.so timeout 300 python tools/layer_bench.py --small c64_64 u128_64 >> gpurun_out/exp1.log 2>&1
done
cat gpurun_out/exp1.log | grep -v Warn | cut -c1-120
