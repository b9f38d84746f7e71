#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_v5_b32.json 2> gpurun_out/bench_v5_b32.err
python bench.py > gpurun_out/bench_v5_b16.json 2> gpurun_out/bench_v5_b16.err
python bench.py --impl reference --steps 2 > gpurun_out/bench_v5_ref.json 2> gpurun_out/bench_v5_ref.err
