#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_frame_metrics.csv python tools/profile_frame.py > gpurun_out/prof_a.log 2>&1; echo "a rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_framemode_metrics.csv python tools/profile_frame.py --mode frame > gpurun_out/prof_b.log 2>&1; echo "b rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 21 -c 3 -f -o gpurun_out/r2_slice2 python tools/profile_frame.py > gpurun_out/prof_c.log 2>&1; echo "c rc=$?"
ls -la gpurun_out/*.ncu-rep
