#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "uint8 or golden" 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-side > gpurun_out/bench_u8_b16.json 2> gpurun_out/bench_u8.err; tail -3 gpurun_out/bench_u8.err
python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_u8_b32.json 2>> gpurun_out/bench_u8.err
