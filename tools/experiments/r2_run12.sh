#!/bin/bash
# round 2, call 12: profile of the current build (launch list, ncu --set full of the GEMM instantiations + attention,
# gemm_traffic.json with the build id), guard tests, head microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_guards.py tests/test_gpu_trainers.py -m gpu -q > gpurun_out/r2_k_pytest.log 2>&1; echo "guards+trainers rc=$?"; tail -3 gpurun_out/r2_k_pytest.log
bash tools/run_profile_r2.sh r2a 2>&1 | tail -40
timeout 600 python tools/microbench_heads.py > gpurun_out/r2_k_microbench_heads.jsonl 2> gpurun_out/r2_k_microbench_heads.err; echo "heads rc=$?"
