#!/bin/bash
# Round-2 profile of the CURRENT build (default operand dtype fp16, bf16 beside it): launch list, ncu --set full of every
# GEMM instantiation and the attention kernel of a step, derived profiles/gemm_traffic.json (stamped with the build id).
# usage (GPU box): bash tools/run_profile_r2.sh <tag>     outputs under gpurun_out/
tag=${1:-r2}
BID=$(python -c "from eoe_b200 import _lib; print(_lib.lib().eoe_build_id().decode())")
echo "build id $BID"
for dt in f16 bf16; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --dtype $dt"
  timeout 300 $CMD > gpurun_out/plain_${tag}_$dt.log 2>&1 || { echo "plain run failed ($dt)"; continue; }
  if [ $dt = f16 ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv --log-file gpurun_out/launches_${tag}_$dt.csv $CMD > gpurun_out/ncu1_${tag}_$dt.log 2>&1
    echo "launch list rc=$?"
    python tools/summarize_launches.py gpurun_out/launches_${tag}_$dt.csv "Launch list, build $BID ($dt), ViT-B/16, batch 512" "ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 $CMD" > gpurun_out/${tag}_launches_vitb16_b512_$dt.md
  fi
  # one full block of the third step: QKV, attention, out-proj, c_fc, c_proj (+ patch embed earlier)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_tc" -s 50 -c 7 -o gpurun_out/prof_${tag}_$dt $CMD > gpurun_out/ncu2_${tag}_$dt.log 2>&1
  echo "full capture rc=$? ($dt)"
  python tools/summarize_ncu_full.py gpurun_out/prof_${tag}_$dt.ncu-rep "ncu --set full, build $BID ($dt): GEMM instantiations + attention of one step" "ncu --set full --clock-control none --import-source on -k regex:gemm_kernel|attention_tc -s 50 -c 7 $CMD" > gpurun_out/${tag}_ncu_full_$dt.md
done
python tools/make_gemm_traffic.py $BID f16=gpurun_out/prof_${tag}_f16.ncu-rep bf16=gpurun_out/prof_${tag}_bf16.ncu-rep > gpurun_out/gemm_traffic_$tag.json
cat gpurun_out/gemm_traffic_$tag.json
# gpurun copies back at most 64 MiB: the summaries above are the evidence, the raw reports stay on the box
rm -f gpurun_out/prof_${tag}_bf16.ncu-rep
[ $(stat -c %s gpurun_out/prof_${tag}_f16.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ] && rm -f gpurun_out/prof_${tag}_f16.ncu-rep
ls -la gpurun_out | head -30
