#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_r_bench_settle_$i.json 2> gpurun_out/r2_r_bench_settle_$i.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_r_bench_settle_$i.json'))
print(round(d['value']), d['ms_per_step'], d['clocks'])
print('e2e', round(d['e2e']['value']), d['e2e']['timing'], d['e2e']['clocks'])
print('f32', round(d['e2e_f32']['value']), 'raw', round(d['e2e_raw']['value']), d['side_metrics']['other_dtype']['value'], d['roofline']['frac'], d['roofline']['traffic'])
"
done
timeout 600 python bench.py --no-cpu-baseline --no-side --settle-s 0 > gpurun_out/r2_r_bench_nosettle.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_r_bench_nosettle.json'))
print('no settle', round(d['value']), d['clocks'], 'e2e', round(d['e2e']['value']), d['e2e']['clocks'])
"
