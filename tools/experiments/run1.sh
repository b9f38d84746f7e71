set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_gpu_heads.py tests/test_gpu_auc.py -m gpu -x -q 2>&1 | tail -25
python tools/microbench_heads.py > gpurun_out/microbench_heads_r1.jsonl 2> gpurun_out/microbench_heads_r1.err; tail -5 gpurun_out/microbench_heads_r1.err; cat gpurun_out/microbench_heads_r1.jsonl
