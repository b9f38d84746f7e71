#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "resize or raw or uint8" 2>&1 | tail -12
