#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_heads.py tests/test_gpu_trainers.py -m gpu -x -q 2>&1 | tail -15
