#!/bin/bash
# round-end validation of the build that ships: full GPU suite, smoke(), default bench line, head / AUC microbench, latencies
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest_all.log | cut -c1-200
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2t_smoke.log
python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; echo "bench rc=$?"
python tools/microbench_heads.py > gpurun_out/r2t_microbench_heads.jsonl 2> gpurun_out/r2t_microbench_heads.err; echo "microbench rc=$?"

python -c "
import json
d=json.load(open('gpurun_out/r2t_bench_default.json'))
print(round(d['value']), round(d['e2e']['value']), d['roofline']['traffic'], round(d['roofline']['frac'],3), d['clocks'], d['side_metrics']['auc'])
"
grep '"auc' gpurun_out/r2t_microbench_heads.jsonl | cut -c1-160
