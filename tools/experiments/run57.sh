#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
python tools/timeline.py --batch 512 --steps 3 > gpurun_out/timeline_v8.md 2>/dev/null
