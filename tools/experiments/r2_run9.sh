#!/bin/bash
# round 2, call 9: single-launch AUC (templated), graph step, and the compute-sanitizer passes SURVEY section 5 promised
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py -m gpu -q -x > gpurun_out/r2_h_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -2 gpurun_out/r2_h_pytest_auc.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_h_latency.err > gpurun_out/r2_h_latency.jsonl; grep '"auc"' gpurun_out/r2_h_latency.jsonl | cut -c1-330
SAN="compute-sanitizer --error-exitcode 99 --print-limit 20"
SEL_HEADS='tests/test_gpu_heads.py -k golden'
SEL_AUC='tests/test_gpu_auc.py -k "golden or (single_launch and 3000)"'
SEL_ENC='tests/test_gpu_encoder.py -k "(gemm_vs_torch_fp32 and 257-768-768) or (gemm_lnfold and 100-256-768) or (gemm_residual_stats and 100-768-768) or (attention_vs_torch and (3-197 or 2-50)) or single_pass_rescale"'
for tool in memcheck racecheck; do
  for sel in heads auc enc; do
    case $sel in heads) S="$SEL_HEADS";; auc) S="$SEL_AUC";; enc) S="$SEL_ENC";; esac
    eval timeout 900 $SAN --tool $tool python -m pytest $S -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_h_sanitizer_${tool}_${sel}.log 2>&1
    echo "$tool $sel rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_h_sanitizer_${tool}_${sel}.log | tail -3
  done
done
