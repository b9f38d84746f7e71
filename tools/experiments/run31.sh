#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1_v4.json 2> gpurun_out/bench_r1_v4.err
tail -2 gpurun_out/bench_r1_v4.err
