#!/bin/bash
# Re-entry probe (one B200): full GPU suite on HEAD, then ncu --set full of the low-K launches of the frame
# (conv4_1, the KernelFilter down / up convolutions, the first 1x1 shortcut) for the stall analysis.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_all.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 7 -c 8 -f -o gpurun_out/r2b_lowk python tools/profile_frame.py > gpurun_out/prof_lowk.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/layer_bench.py --small c64_512 s512_256 s256_128 s128_64 c512_64 2>&1 | tail -8
ls -la gpurun_out/*.ncu-rep
