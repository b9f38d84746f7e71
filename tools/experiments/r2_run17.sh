#!/bin/bash
mkdir -p gpurun_out
for b in 512 1024 512 1024 768; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --batch $b > gpurun_out/r2_p_bench_b$b.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2_p_bench_b$b.json'))
print($b, round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], {k:round(v) for k,v in d['roofline']['per_kind_tflops'].items()})
"
done
