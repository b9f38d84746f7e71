#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu -k "attention or encoder_vs_oracle or batching or uint8" > gpurun_out/t65.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t65.log
for i in 1 2; do timeout 100 python tools/attn_small_probe.py 0; timeout 100 python tools/attn_small_probe.py 128; done
timeout 300 python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/b32_65.json 2> gpurun_out/b32_65.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/b32_65.json"))
print(d["config"]["batch_per_gpu"], round(d["value"]), round(d["e2e_u8"]["value"]), d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
PY
