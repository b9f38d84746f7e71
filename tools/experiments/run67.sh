#!/bin/bash
# ncu --set full of the two tcgen05 attention kernels of the final build (L = 197 with early refill, L <= 64 pair kernel)
CMD16="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side"
CMD32="python bench.py --patch 32 --prompts 10 --steps 2 --warmup 3 --no-cpu-baseline --no-side"
timeout 300 $CMD16 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc" -s 20 -c 2 -f -o gpurun_out/prof_attn197_v12 $CMD16 > gpurun_out/ncu_attn197.log 2>&1; echo rc=$?
timeout 300 $CMD32 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc64" -s 20 -c 2 -f -o gpurun_out/prof_attn64_v12 $CMD32 > gpurun_out/ncu_attn64.log 2>&1; echo rc=$?
cuobjdump -sass eoe_b200/libeoe_b200.so 2>/dev/null | grep -E "UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|STTM" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head
