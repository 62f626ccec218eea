#!/bin/bash
# compute-sanitizer passes over small end-to-end runs (global mode incl. the staged TMA stores, frame mode, warp, Vgg19 backward)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  echo "===== $tool" >> gpurun_out/sanitizer.log
  timeout 1200 $S --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small_q3-tc or full_epilogue_chain or head_fused or frame_mode_matches or temporal_loss_and_backward or fused_statistics or vgg19_loss" >> gpurun_out/sanitizer.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer.log
done
grep -n "=====\|ERROR SUMMARY\|passed\|failed\|rc=\|RACECHECK SUMMARY\|hazard" gpurun_out/sanitizer.log | head -40
