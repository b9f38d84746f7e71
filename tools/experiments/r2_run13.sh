#!/bin/bash
# round 2, call 13 (8 GPUs): dp_bench + bench at N = 8
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dp_bench.py > gpurun_out/r2_l_dp_bench_n8.jsonl 2> gpurun_out/r2_l_dp_bench_n8.err; echo "dp_bench rc=$?"; cat gpurun_out/r2_l_dp_bench_n8.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_l_bench_n8.json 2> gpurun_out/r2_l_bench_n8.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_l_bench_n8.json; tail -3 gpurun_out/r2_l_bench_n8.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/dp_check.py > gpurun_out/r2_l_dp_check_n8.json 2>> gpurun_out/r2_l_dp_bench_n8.err; echo "dp_check rc=$?"; cat gpurun_out/r2_l_dp_check_n8.json
