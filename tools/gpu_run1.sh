#!/bin/bash
# GPU box script: head-pair smoke first (falls back to RRV_HEAD_PAIR=0), full GPU tests, bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "rgb_head or head_fused or non_multiple" > gpurun_out/t_head.log 2>&1
rc=$?; echo "head rc=$rc" | tee -a gpurun_out/t_head.log
if [ $rc -ne 0 ]; then export RRV_HEAD_PAIR=0; echo "HEAD_PAIR disabled" | tee -a gpurun_out/t_head.log;
  timeout 600 python -m pytest tests -m gpu -q -x -k "rgb_head or head_fused or non_multiple" > gpurun_out/t_head_nopair.log 2>&1; echo "rc=$?" >> gpurun_out/t_head_nopair.log; fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/bench200.json 2> gpurun_out/bench200.err; echo "bench200 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref2.json 2> gpurun_out/ref2.err; echo "ref rc=$?"
tail -5 gpurun_out/t_all.log
