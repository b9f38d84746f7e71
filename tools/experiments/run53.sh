#!/bin/bash
timeout 300 python tools/gemm_probe.py 4 > gpurun_out/gemm_probe4.json 2> gpurun_out/gemm_probe4.err; tail -2 gpurun_out/gemm_probe4.err
