#!/bin/bash
# 8-GPU box: the weak-scaling bench line at N = 8 and 4, BASELINE configs[3] as written (tools/config4.py), and config 3's clip on one GPU.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "bench n4 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/config4.py > gpurun_out/n8_config4.json 2> gpurun_out/n8_config4.err; echo "config4 n8 rc=$?"
cat gpurun_out/n8_config4.json
timeout 600 python tools/config4.py --frames 256 > gpurun_out/n1_config3_clip.json 2> gpurun_out/n1_config3_clip.err; echo "clip256 n1 rc=$?"
cat gpurun_out/n1_config3_clip.json
