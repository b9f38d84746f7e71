#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
python tools/ab_encoder.py --debug-flags pdl=0,nopdl=32 > gpurun_out/ab_pdl.json 2> gpurun_out/ab_pdl.err; tail -2 gpurun_out/ab_pdl.err
