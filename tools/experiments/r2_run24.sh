#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_contract.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 --no-side > gpurun_out/r2_w_bench_n2.json 2> gpurun_out/r2_w_bench_n2.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_w_bench_n2.json'))
print(round(d['value']), d['ms_per_step'], d['per_rank_scoring_ms_per_step'], d['clocks'])
print('e2e', round(d['e2e']['value']), d['e2e']['timing'])
"
