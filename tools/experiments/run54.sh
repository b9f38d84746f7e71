#!/bin/bash
timeout 300 python tools/gemm_probe.py 0,$((66*256)),16,1,$((66*256+1)),17 > gpurun_out/gemm_probe5.json 2> gpurun_out/gemm_probe5.err; tail -2 gpurun_out/gemm_probe5.err
