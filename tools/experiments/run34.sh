#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "attention" 2>&1 | tail -4
timeout 300 python tools/attn_probe.py 512 > gpurun_out/attn_probe.json 2> gpurun_out/attn_probe.err; tail -3 gpurun_out/attn_probe.err
