#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py tests/test_gpu_trainers.py -m gpu -q > gpurun_out/r2_e_pytest.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_e_pytest.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_e_latency.err | grep '"auc"\|train_cls' > gpurun_out/r2_e_latency.jsonl; cat gpurun_out/r2_e_latency.jsonl
