#!/bin/bash
# key-split attention: parity, then in-process A/B (debug bit 7 = single-pass form), then the head-kernel ncu pass
timeout 300 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu -k "attention or encoder_vs_oracle or batching" > gpurun_out/t62.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t62.log
timeout 200 python tools/attn_probe.py > gpurun_out/attn_probe_62.json 2> gpurun_out/attn_probe_62.err; tail -c 600 gpurun_out/attn_probe_62.json
timeout 300 python tools/ab_encoder.py --debug-flags split=0,single=128 --rounds 5 --block 10 > gpurun_out/ab62.json 2>> gpurun_out/ab62.err; python - <<'PY'
import json; d=json.load(open("gpurun_out/ab62.json")); print({k:(v["images_per_s"] if isinstance(v,dict) else v) for k,v in d.items()})
PY
timeout 120 python tools/heads_once.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"clip_score_mma|clip_oe_loss_mma|hsc_rows_kernel|bce_kernel|auc_sort_pass" -c 28 -o gpurun_out/prof_heads_v10 -f python tools/heads_once.py > gpurun_out/ncu3_v10.log 2>&1
echo heads rc=$?
