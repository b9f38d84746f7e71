#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --no-side > gpurun_out/r2_u_bench_n2.json 2> gpurun_out/r2_u_bench_n2.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_u_bench_n2.json'))
print(round(d['value']), d['ms_per_step'], d['per_rank_scoring_ms_per_step'], d['clocks'])
print('e2e', round(d['e2e']['value']), d['e2e']['timing'])
"
python bench.py --no-cpu-baseline --no-side --patch 32 --prompts 10 > gpurun_out/r2_u_bench_b32.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_u_bench_b32.json'))
print('B/32', round(d['value']), d['per_rank_scoring_ms_per_step'], 'e2e', round(d['e2e']['value']), 'f32', round(d['e2e_f32']['value']), 'raw', round(d['e2e_raw']['value']), d['roofline']['frac'])
"
