python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -30
