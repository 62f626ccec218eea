#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conv or golden or facade or non_multiple" > gpurun_out/t_new.log 2>&1; echo "new rc=$?" | tee -a gpurun_out/t_new.log
tail -15 gpurun_out/t_new.log
timeout 300 python tools/layer_bench.py --small --full c64_64 > gpurun_out/lb.log 2>&1
timeout 300 python tools/layer_bench.py --small c64_64 u128_64 >> gpurun_out/lb.log 2>&1
RRV_NO_OSTAGE=1 timeout 300 python tools/layer_bench.py --small --full c64_64 >> gpurun_out/lb.log 2>&1
RRV_NO_OSTAGE=1 timeout 300 python tools/layer_bench.py --small c64_64 u128_64 >> gpurun_out/lb.log 2>&1
grep -v Warn gpurun_out/lb.log | cut -c1-100
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench20 rc=$?"
