#!/bin/bash
# round 2, call 4: single-launch AUC with ballot ranking, graph-captured step fixes, latency microbench, dp_bench at N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py tests/test_gpu_trainers.py -m gpu -q -x > gpurun_out/r2_d_pytest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_d_pytest.log
timeout 600 python tools/microbench_latency.py > gpurun_out/r2_d_latency.jsonl 2> gpurun_out/r2_d_latency.err; echo "latency rc=$?"; cat gpurun_out/r2_d_latency.jsonl; tail -3 gpurun_out/r2_d_latency.err | cut -c1-300
timeout 600 python tools/dp_bench.py > gpurun_out/r2_d_dp_bench_n1.jsonl 2> gpurun_out/r2_d_dp_bench_n1.err; echo "dp_bench rc=$?"; cat gpurun_out/r2_d_dp_bench_n1.jsonl; tail -3 gpurun_out/r2_d_dp_bench_n1.err | cut -c1-300
