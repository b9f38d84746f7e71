#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/experiments/r2_graph_debug.py > gpurun_out/r2_f_graph_debug.log 2>&1; tail -40 gpurun_out/r2_f_graph_debug.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_trainers.py -m gpu -q -k graph > gpurun_out/r2_f_pytest.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_f_pytest.log; grep -n "^E " gpurun_out/r2_f_pytest.log | head
timeout 900 python -m pytest tests/test_gpu_auc.py -m gpu -q > gpurun_out/r2_f_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2_f_pytest_auc.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_f_latency.err | grep '"auc"' > gpurun_out/r2_f_latency.jsonl; cut -c1-420 gpurun_out/r2_f_latency.jsonl
