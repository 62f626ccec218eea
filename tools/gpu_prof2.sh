#!/bin/bash
# Round-2 profile set (one B200): launch list of bench.py, per-launch metrics of one frame (global / frame mode),
# ncu --set full of the slice2 kernels + the RGB head and of a 256->256 layer, kernel-only times of the warp kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bf16 --no-side > gpurun_out/prof_l.log 2>&1; echo "l rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_frame_metrics.csv python tools/profile_frame.py > gpurun_out/prof_a.log 2>&1; echo "a rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_framemode_metrics.csv python tools/profile_frame.py --mode frame > gpurun_out/prof_b.log 2>&1; echo "b rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 21 -c 3 -f -o gpurun_out/r2_slice2 python tools/profile_frame.py > gpurun_out/prof_c.log 2>&1; echo "c rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 5 -c 1 -f -o gpurun_out/r2_conv3_3 python tools/profile_frame.py > gpurun_out/prof_d.log 2>&1; echo "d rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:warp_ -c 12 --csv --log-file gpurun_out/r2_warp_kernels.csv python tools/temporal_bench.py > gpurun_out/prof_e.log 2>&1; echo "e rc=$?"
ls -la gpurun_out/*.ncu-rep
