#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
python tools/ab_encoder.py --variants fold > gpurun_out/ab_cur.json 2> gpurun_out/ab_cur.err
EOE_B200_LIB=$PWD/tools/_variants/libeoe_b200_v3.so python tools/ab_encoder.py --variants fold > gpurun_out/ab_v3.json 2> gpurun_out/ab_v3.err
python tools/ab_encoder.py --variants fold > gpurun_out/ab_cur_2.json 2> gpurun_out/ab_cur_2.err
python tools/timeline.py --batch 512 --steps 3 > gpurun_out/timeline_v5.md 2>/dev/null
