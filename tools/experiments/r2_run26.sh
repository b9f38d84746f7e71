#!/bin/bash
# final-build evidence: launch list + ncu --set full of the GEMMs / attention (f16, bf16), gemm_traffic.json stamped with this
# build id, the default bench line quoting it, heads microbench, ncu --set full of the tcgen05 CLIP heads
mkdir -p gpurun_out
bash tools/run_profile_r2.sh r2k > gpurun_out/r2k_profile.log 2>&1; echo "profile rc=$?"
if grep -q dram_bytes_per_launch gpurun_out/gemm_traffic_r2k.json; then cp gpurun_out/gemm_traffic_r2k.json profiles/gemm_traffic.json; fi
python bench.py > gpurun_out/r2k_bench_final.json 2> gpurun_out/r2k_bench_final.err; echo "bench rc=$?"
python tools/microbench_heads.py > gpurun_out/r2k_microbench_heads.jsonl 2> gpurun_out/r2k_microbench_heads.err; echo "microbench rc=$?"
CMD="python tools/experiments/clip_loss_tc_probe.py"
timeout 120 $CMD > gpurun_out/r2k_clip_probe_plain.log 2>&1; echo "probe rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clip_oe_loss_tc -s 3 -c 1 -o gpurun_out/prof_r2k_clip_loss $CMD > gpurun_out/r2k_ncu_clip_loss.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu_full.py gpurun_out/prof_r2k_clip_loss.ncu-rep "ncu --set full of cliptc::clip_oe_loss_tc_kernel<bf16>, final layout (1 M rows, d = 512, K = 30)" "ncu --set full --clock-control none --import-source on -k regex:clip_oe_loss_tc -s 3 -c 1 $CMD" > gpurun_out/r2k_ncu_clip_loss_tc.md
ncu -i gpurun_out/prof_r2k_clip_loss.ncu-rep --page raw --csv > gpurun_out/r2k_ncu_clip_loss_raw.csv 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_final.json'))
print(round(d['value']), round(d['e2e']['value']), d['roofline']['traffic'], round(d['roofline']['frac'],3), d['clocks'])
"
tail -4 gpurun_out/r2k_microbench_heads.jsonl | cut -c1-200
