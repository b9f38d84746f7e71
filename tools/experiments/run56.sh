#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
python tools/ab_encoder.py --variants fold --rounds 6 > gpurun_out/ab_tma_cur_$i.json 2>/dev/null
EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_head.so python tools/ab_encoder.py --variants fold --rounds 6 > gpurun_out/ab_tma_head_$i.json 2>/dev/null
done
