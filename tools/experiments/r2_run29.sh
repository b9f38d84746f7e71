#!/bin/bash
# tiled AUC, third step (sort tile size by n, monotone workspace size): parity incl. the new tile-switch tests, A/B of the
# scan tile size (4 / 8 / 16 elements per thread)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_auc.py tests/test_gpu_guards.py -m gpu -x -q > gpurun_out/r2q_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2q_pytest_auc.log | cut -c1-300
S="65536 262144 655360 1000000 4194304 16777216"
timeout 300 python tools/experiments/auc_tiled_probe.py $S > gpurun_out/r2q_auc_probe.jsonl 2> gpurun_out/r2q_auc_probe.err; echo "probe rc=$?"
for v in scan16 scan4; do
  EOE_B200_LIB=tools/_variants/libeoe_b200_$v.so timeout 300 python tools/experiments/auc_tiled_probe.py $S >> gpurun_out/r2q_auc_probe.jsonl 2>> gpurun_out/r2q_auc_probe.err
done
python - <<P
import json
for l in open("gpurun_out/r2q_auc_probe.jsonl"):
    d=json.loads(l); print(d["lib"][-12:], d["n"], d["bit_exact_vs_sklearn"], round(d["auc"]["ms"],4), round(d["auc+ap"]["ms"],4), round(d["auc_f16ties"]["ms"],4))
P
