#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "gemm" 2>&1 | tail -5
timeout 300 python tools/gemm_probe.py 0,16,1,17 > gpurun_out/gemm_probe3.json 2> gpurun_out/gemm_probe3.err; tail -2 gpurun_out/gemm_probe3.err
