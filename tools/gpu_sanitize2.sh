#!/bin/bash
# compute-sanitizer memcheck / synccheck over the late round-2 kernels: staged residual epilogue (cp.async + TMA store), 64-byte
# operand rows, merged hi|lo weight planes of the RGB head, L2 prefetch of the next tile's residual.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
rm -f gpurun_out/sanitizer2.log
for tool in memcheck synccheck; do
  echo "===== $tool" >> gpurun_out/sanitizer2.log
  timeout 1200 $S --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernelfilter_upsample or conv_matches_torch_cpu or small_q3-tc or full_epilogue_chain or head_fused or fold_filter or frame_mode_matches" >> gpurun_out/sanitizer2.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer2.log
done
grep -n "=====\|ERROR SUMMARY\|passed\|failed\|rc=" gpurun_out/sanitizer2.log | head -20
