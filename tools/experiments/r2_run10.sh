#!/bin/bash
# round 2, call 10 (2 GPUs): NCCL world-2 tests, dp_bench, dp_check, bench at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > gpurun_out/r2_i_pytest_dp.log 2>&1; echo "dp tests rc=$?"; grep -E "DP_VS_GLOBAL|passed|failed" gpurun_out/r2_i_pytest_dp.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_bench.py > gpurun_out/r2_i_dp_bench_n2.jsonl 2> gpurun_out/r2_i_dp_bench_n2.err; echo "dp_bench rc=$?"; cat gpurun_out/r2_i_dp_bench_n2.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_bench.py --bucket-mb 8 > gpurun_out/r2_i_dp_bench_n2_b8.jsonl 2>> gpurun_out/r2_i_dp_bench_n2.err; cat gpurun_out/r2_i_dp_bench_n2_b8.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py > gpurun_out/r2_i_dp_check_n2.json 2>> gpurun_out/r2_i_dp_bench_n2.err; echo "dp_check rc=$?"; cat gpurun_out/r2_i_dp_check_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_i_bench_n2.json 2> gpurun_out/r2_i_bench_n2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_i_bench_n2.json; tail -3 gpurun_out/r2_i_bench_n2.err | cut -c1-300
