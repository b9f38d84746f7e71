#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py tests/test_gpu_guards.py -m gpu -q -x -k "auc" > gpurun_out/r2_m_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -2 gpurun_out/r2_m_pytest_auc.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_m_latency.err > gpurun_out/r2_m_latency.jsonl; grep '"auc"' gpurun_out/r2_m_latency.jsonl | cut -c1-330
