#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py -m gpu -q -x > gpurun_out/r2_g_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2_g_pytest_auc.log
timeout 900 python -m pytest tests/test_gpu_trainers.py -m gpu -q -k graph > gpurun_out/r2_g_pytest.log 2>&1; echo "graph tests rc=$?"; tail -3 gpurun_out/r2_g_pytest.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_g_latency.err > gpurun_out/r2_g_latency.jsonl; grep '"auc"' gpurun_out/r2_g_latency.jsonl | cut -c1-330
