#!/bin/bash
# round 2, call 16 (4 GPUs): NCCL AVG + tail bucket check (dp_check), dp_bench, bench at N = 4
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py > gpurun_out/r2_o_dp_check_n4.json 2> gpurun_out/r2_o_n4.err; echo "dp_check rc=$?"; cat gpurun_out/r2_o_dp_check_n4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/dp_bench.py > gpurun_out/r2_o_dp_bench_n4.jsonl 2>> gpurun_out/r2_o_n4.err; echo "dp_bench rc=$?"; cat gpurun_out/r2_o_dp_bench_n4.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_o_bench_n4.json 2>> gpurun_out/r2_o_n4.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_o_bench_n4.json
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "DP_VS_GLOBAL|passed|failed" | grep -v print | cut -c1-300
