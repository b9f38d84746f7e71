#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_t_bench_n8.json 2> gpurun_out/r2_t_bench_n8.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_t_bench_n8.json'))
print(round(d['value']), d['ms_per_step'], d['clocks'])
print('e2e', round(d['e2e']['value']), d['e2e']['timing'], d['e2e']['clocks'])
print('f32', round(d['e2e_f32']['value']), 'raw', round(d['e2e_raw']['value']), d['side_metrics']['other_dtype']['value'])
for r in d['side_metrics']['dp_train']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k!='config'})
"
