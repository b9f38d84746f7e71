#!/bin/bash
# Profile of the PRECISE mode (operand dtype f16x2) on the current build: launch list of one step and ncu --set full of the
# split GEMM instantiations + the split attention kernel.   usage (GPU box): bash tools/run_profile_split.sh <tag>
tag=${1:-r2_split}
BID=$(python -c "from eoe_b200 import _lib; print(_lib.lib().eoe_build_id().decode())")
echo "build id $BID"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --dtype f16x2 --settle-s 0"
timeout 300 $CMD > gpurun_out/plain_${tag}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 140 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu1_${tag}.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${tag}.csv "Launch list, build $BID (f16x2, precise mode), ViT-B/16, batch 512" "ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 140 $CMD" > gpurun_out/${tag}_launches_vitb16_b512.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_tc" -s 50 -c 7 -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "full capture rc=$?"
python tools/summarize_ncu_full.py gpurun_out/prof_${tag}.ncu-rep "ncu --set full, build $BID (f16x2): split GEMM instantiations + split attention of one step" "ncu --set full --clock-control none --import-source on -k regex:gemm_kernel|attention_tc -s 50 -c 7 $CMD" > gpurun_out/${tag}_ncu_full.md
rm -f gpurun_out/prof_${tag}.ncu-rep      # gpurun copies back at most 64 MiB; the summary above is the evidence
