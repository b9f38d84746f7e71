#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_text.py -x -q -m gpu > gpurun_out/t64.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t64.log
timeout 300 python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/b32_64.json 2> gpurun_out/b32_64.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/b32_64.json"))
print(d["config"]["batch_per_gpu"], round(d["value"]), round(d["e2e_u8"]["value"]), d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
PY
