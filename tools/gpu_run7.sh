#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_framemode_metrics.csv python tools/profile_frame.py --mode frame > gpurun_out/prof_b.log 2>&1; echo "b rc=$?"
timeout 600 python tools/frame_mode_bench.py 2>&1 | tail -1
timeout 600 python tools/frame_mode_bench.py 720p 2>&1 | tail -1
