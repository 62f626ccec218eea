#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
D=rerevst-code_b200/csrc
for v in "" _nores; do
  echo "== variant '$v'" >> gpurun_out/exp2.log
  RRV_LIB_PATH=$PWD/$D/librerevst_b200$v.so timeout 300 python tools/layer_bench.py --small --full c64_64 >> gpurun_out/exp2.log 2>&1
done
grep -v Warn gpurun_out/exp2.log | cut -c1-100
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_framemode_metrics.csv python tools/profile_frame.py --mode frame > gpurun_out/prof_b.log 2>&1; echo "b rc=$?"
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_stat or frame_mode or pointwise" 2>&1 | tail -3
