#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-side --steps 200 > gpurun_out/r2_v_bench_k200.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_v_bench_k200.json'))
print('K=200', round(d['value']), d['ms_per_step'], d['per_rank_scoring_ms_per_step'], d['clocks'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['encoder_tensor_frac'])
"
python bench.py --no-cpu-baseline --no-side --steps 200 --dtype bf16 > gpurun_out/r2_v_bench_k200_bf16.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_v_bench_k200_bf16.json'))
print('K=200 bf16', round(d['value']), d['ms_per_step'], d['clocks'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['encoder_tensor_frac'])
"
