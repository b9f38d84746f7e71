#!/bin/bash
# L = 197 attention with per-buffer load barriers / early refill: parity, stand-alone timing vs the previous build, step A/B
timeout 300 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu -k "attention or encoder_vs_oracle or batching or full_batch" > gpurun_out/t66.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t66.log
for i in 1 2; do
  EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_head.so timeout 100 python tools/attn_probe.py | tr -d '\n '; echo " (head)"
  timeout 100 python tools/attn_probe.py | tr -d '\n '; echo " (new)"
done
for r in 1 2 3; do
  for v in head new; do
    if [ $v = head ]; then export EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_head.so; else unset EOE_B200_LIB; fi
    timeout 200 python tools/ab_encoder.py --variants fold --rounds 3 --block 10 > gpurun_out/ab66_${v}_$r.json 2>> gpurun_out/ab66.err
    python - <<PY
import json; d=json.load(open("gpurun_out/ab66_${v}_$r.json")); print("$v $r", d["fold"]["images_per_s"])
PY
  done
done
