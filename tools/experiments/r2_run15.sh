#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_auc.py -m gpu -q -x > gpurun_out/r2_n_pytest_auc.log 2>&1; echo "auc tests rc=$?"; tail -4 gpurun_out/r2_n_pytest_auc.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_guards.py -m gpu -q -k auc > gpurun_out/r2_n_pytest_guards.log 2>&1; echo "guards rc=$?"; tail -2 gpurun_out/r2_n_pytest_guards.log
timeout 600 python tools/microbench_latency.py 2> gpurun_out/r2_n_latency.err > gpurun_out/r2_n_latency.jsonl; grep '"auc"' gpurun_out/r2_n_latency.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.items() if k in ('n','us','launches','device_us_graph_replay','one_cta_device_us_graph_replay','tiled_device_us_graph_replay','phase_clocks_keys_sort__scans__terms_sum','sklearn_host_us')})
"
