#!/bin/bash
# flags: 1 = main loop only; (pairs << 8)
python tools/gemm_probe.py 1,$((37*256+1)),$((18*256+1)),$((8*256+1)),0,$((37*256)) > gpurun_out/gemm_probe2.json 2> gpurun_out/gemm_probe2.err
nvidia-smi -q -d POWER | grep -i -E "power limit|draw" | head -8 > gpurun_out/power.txt
