#!/bin/bash
# round 2, call 18: full GPU suite on the final build, profile (launch list, ncu --set full, gemm_traffic.json), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_q_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/r2_q_pytest_all.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_q_smoke.log
bash tools/run_profile_r2.sh r2b 2>&1 | grep -E "build id|rc=" 
cp gpurun_out/gemm_traffic_r2b.json profiles/gemm_traffic.json
timeout 900 python bench.py > gpurun_out/r2_q_bench_default.json 2> gpurun_out/r2_q_bench_default.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_q_bench_default.json'))
print(d['value'], d['e2e']['value'], d['roofline']['traffic'], d['roofline']['traffic_of'])
"
