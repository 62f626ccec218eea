#!/bin/bash
# First-layer kernel: parity tests, kernel time (ncu, one frame) and a 20-step bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "first_layer or global_mode or facade" 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_write.sum -k regex:first_layer --csv --log-file gpurun_out/fl_time.csv python tools/profile_frame.py > gpurun_out/fl_time.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/fl_time.csv | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-side 2>/dev/null | python tools/benchline.py
