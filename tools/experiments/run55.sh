#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
python tools/ab_encoder.py --debug-flags gelux=0,plain=64 --rounds 8 > gpurun_out/ab_gelux.json 2> gpurun_out/ab_gelux.err; tail -2 gpurun_out/ab_gelux.err
