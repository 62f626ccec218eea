#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_stat or fold_filter or frame_mode or prepass or temporal_loss_config5 or golden" > gpurun_out/t_new.log 2>&1; echo "new rc=$?" | tee -a gpurun_out/t_new.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench20 rc=$?"
tail -5 gpurun_out/t_new.log; tail -5 gpurun_out/t_all.log
