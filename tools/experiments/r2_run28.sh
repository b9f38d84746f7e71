#!/bin/bash
# tiled AUC, second step (pairwise tree in shared memory, keys kernel with one large block per SM, info folded into the tree
# kernel): parity, A/B against variants (previous library, 2048-key sort tiles, ballot ranking, round-1 keys kernel),
# launch list and ncu --set full of one sort pass at 1 M scores
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_auc.py tests/test_gpu_guards.py -m gpu -x -q > gpurun_out/r2o_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -3 gpurun_out/r2o_pytest_auc.log | cut -c1-300
timeout 300 python tools/experiments/auc_tiled_probe.py > gpurun_out/r2o_auc_probe.jsonl 2> gpurun_out/r2o_auc_probe.err; echo "probe rc=$?"
for v in aucold s8 ballot keys1; do
  EOE_B200_LIB=tools/_variants/libeoe_b200_$v.so timeout 300 python tools/experiments/auc_tiled_probe.py >> gpurun_out/r2o_auc_probe.jsonl 2>> gpurun_out/r2o_auc_probe.err
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_auc_1m_launches.csv python tools/auc_small_once.py 1000000 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:auc_sort_pass -s 4 -c 1 -o gpurun_out/prof_r2o_auc_sort python tools/auc_small_once.py 1000000 > gpurun_out/r2o_ncu_sort.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu_full.py gpurun_out/prof_r2o_auc_sort.ncu-rep "ncu --set full of auc_sort_pass_kernel (pass 0 of the second call, 1 M scores)" "ncu --set full --clock-control none --import-source on -k regex:auc_sort_pass -s 4 -c 1 python tools/auc_small_once.py 1000000" > gpurun_out/r2o_ncu_auc_sort.md
ncu -i gpurun_out/prof_r2o_auc_sort.ncu-rep --page details --csv > gpurun_out/r2o_ncu_auc_sort_details.csv 2>/dev/null
python - <<P
import csv, json
rows=list(csv.reader(l for l in open("gpurun_out/r2o_auc_1m_launches.csv") if l.startswith('"')))
h=rows[0]
for r in rows[-10:]:
    print(r[h.index("Kernel Name")][:40], r[-1])
for l in open("gpurun_out/r2o_auc_probe.jsonl"):
    d=json.loads(l); print(d["lib"][-12:], d["n"], d["bit_exact_vs_sklearn"], round(d["auc"]["ms"],4), round(d["auc+ap"]["ms"],4), round(d["auc_f16ties"]["ms"],4))
P
