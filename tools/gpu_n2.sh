#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_dist.py -m gpu -q -x > gpurun_out/n2_dist.log 2>&1; echo "dist rc=$?" | tee -a gpurun_out/n2_dist.log
tail -3 gpurun_out/n2_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/n2_bench.err
