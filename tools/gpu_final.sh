#!/bin/bash
# End-of-round set on one B200: GPU tests, smoke, both bench lines (driver-style 20 steps and the 200-step default), the reference
# arm, then the profile set (tools/gpu_prof2.sh).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/bench200.json 2> gpurun_out/bench200.err; echo "bench200 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref20.json 2> gpurun_out/ref20.err; echo "ref rc=$?"
bash tools/gpu_prof2.sh
