#!/bin/bash
python tools/gemm_probe.py > gpurun_out/gemm_probe.json 2> gpurun_out/gemm_probe.err
EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_v3.so python tools/gemm_probe.py 0 > gpurun_out/gemm_probe_v3.json 2>> gpurun_out/gemm_probe.err
