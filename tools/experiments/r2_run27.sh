#!/bin/bash
# tiled AUC rework (sort pass: tile reordered by digit in shared memory before the scatter; distinct / corner scans: warp-striped
# coalesced accesses): parity first, then A/B against the previous library and the direct-scatter variant, then the launch list at 1 M scores
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_auc.py tests/test_gpu_guards.py -m gpu -x -q > gpurun_out/r2n_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2n_pytest_auc.log | cut -c1-300
timeout 300 python tools/experiments/auc_tiled_probe.py > gpurun_out/r2n_auc_probe.jsonl 2> gpurun_out/r2n_auc_probe.err; echo "probe rc=$?"
EOE_B200_LIB=tools/_variants/libeoe_b200_aucold.so timeout 300 python tools/experiments/auc_tiled_probe.py >> gpurun_out/r2n_auc_probe.jsonl 2>> gpurun_out/r2n_auc_probe.err
EOE_B200_LIB=tools/_variants/libeoe_b200_aucdirect.so timeout 300 python tools/experiments/auc_tiled_probe.py >> gpurun_out/r2n_auc_probe.jsonl 2>> gpurun_out/r2n_auc_probe.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_auc_1m_launches.csv python tools/auc_small_once.py 1000000 > /dev/null 2>&1
python - <<P
import csv, json
rows=list(csv.reader(l for l in open("gpurun_out/r2n_auc_1m_launches.csv") if l.startswith('"')))
h=rows[0]
for r in rows[-11:]:
    print(r[h.index("Kernel Name")][:40], r[-1])
for l in open("gpurun_out/r2n_auc_probe.jsonl"):
    d=json.loads(l); print(d["lib"][-12:], d["n"], d["bit_exact_vs_sklearn"], round(d["auc"]["ms"],4), round(d["auc+ap"]["ms"],4), round(d["auc_f16ties"]["ms"],4))
P
