bash tools/run_debug.sh attn attn_bench
python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
python tools/debug_gemm.py attn_bench > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 5 -c 1 -o gpurun_out/prof_attn_r1 python tools/debug_gemm.py attn_bench > gpurun_out/ncu_attn.log 2>&1
echo ncu rc=$?
