#!/bin/bash
# tiled AUC with programmatic dependent launch between its kernels: parity, A/B against the same build without it,
# latency microbench (replays the tiled pipeline from a CUDA graph: PDL edges under capture)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_auc.py tests/test_gpu_guards.py tests/test_gpu_trainers.py -m gpu -x -q > gpurun_out/r2s_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2s_pytest_auc.log | cut -c1-300
S="65536 262144 1000000 4194304 16777216"
timeout 300 python tools/experiments/auc_tiled_probe.py $S > gpurun_out/r2s_auc_probe.jsonl 2> gpurun_out/r2s_auc_probe.err; echo "probe rc=$?"
EOE_B200_LIB=tools/_variants/libeoe_b200_nopdl.so timeout 300 python tools/experiments/auc_tiled_probe.py $S >> gpurun_out/r2s_auc_probe.jsonl 2>> gpurun_out/r2s_auc_probe.err
timeout 150 python tools/microbench_latency.py > gpurun_out/r2s_latency.jsonl 2> gpurun_out/r2s_latency.err; echo "latency rc=$?"
python - <<P
import json
for l in open("gpurun_out/r2s_auc_probe.jsonl"):
    d=json.loads(l); print(d["lib"][-12:], d["n"], d["bit_exact_vs_sklearn"], round(d["auc"]["ms"],4), round(d["auc+ap"]["ms"],4), round(d["auc_f16ties"]["ms"],4))
for l in open("gpurun_out/r2s_latency.jsonl"):
    d=json.loads(l)
    if d.get("kernel")=="auc": print(d["n"], d.get("us"), d.get("device_us_graph_replay"), d.get("tiled_us"), d.get("tiled_device_us_graph_replay"))
P
