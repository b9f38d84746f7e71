#!/bin/bash
# round 2, call 2: f16 as the default operand dtype -- end-to-end score parity on the GPU, single-pass f16 softmax, tanh GELU
# for f16 (accuracy + speed A/B against the exp+divide GELU and the two-pass softmax, as alternate libraries)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder.py -m gpu -q -s -k "attention or end_to_end or golden" > gpurun_out/r2_b_pytest_encoder.log 2>&1
echo "encoder tests rc=$?"
grep END_TO_END gpurun_out/r2_b_pytest_encoder.log
for v in geluexact twopass; do
  EOE_B200_LIB=tools/_variants/libeoe_b200_$v.so python -m pytest tests/test_gpu_encoder.py -m gpu -q -s -k "end_to_end" > gpurun_out/r2_b_pytest_e2e_$v.log 2>&1
  echo "$v rc=$?"; grep END_TO_END gpurun_out/r2_b_pytest_e2e_$v.log
done
for i in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype f16 > gpurun_out/r2_b_bench_f16_default_$i.json 2>/dev/null
  for v in geluexact twopass; do
    EOE_B200_LIB=tools/_variants/libeoe_b200_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype f16 > gpurun_out/r2_b_bench_f16_${v}_$i.json 2>/dev/null
  done
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype bf16 > gpurun_out/r2_b_bench_bf16_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, d['dtype'], round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v) for k,v in d['roofline']['per_kind_tflops'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_b_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/r2_b_pytest_all.log
