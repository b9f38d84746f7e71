#!/bin/bash
# round-1 re-entry: GPU parity of the centred 16-bit residual copy + alternating-process A/B against the HEAD build
python -m pytest tests/test_gpu_encoder.py -x -q -m gpu > gpurun_out/t61_encoder.log 2>&1; echo "pytest encoder rc=$?"; tail -5 gpurun_out/t61_encoder.log
for r in 1 2 3; do
  for v in head new; do
    if [ $v = head ]; then export EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_head.so; else unset EOE_B200_LIB; fi
    python tools/ab_encoder.py --variants fold --rounds 3 --block 10 > gpurun_out/ab61_${v}_$r.json 2>> gpurun_out/ab61.err
    python - <<PY
import json; d=json.load(open("gpurun_out/ab61_${v}_$r.json")); print("$v $r", d["fold"]["images_per_s"], d["fold"]["gemm_us"])
PY
  done
done
