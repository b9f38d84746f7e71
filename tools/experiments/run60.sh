#!/bin/bash
python bench.py --no-cpu-baseline --no-side > gpurun_out/bench_raw_b16.json 2> gpurun_out/bench_raw.err; tail -3 gpurun_out/bench_raw.err
python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_raw_b32.json 2>> gpurun_out/bench_raw.err
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from eoe_b200.encoder import ClipImageEncoder
from eoe_b200.synth import random_vit_state_dict
for patch,(h,w) in ((16,(375,500)),(32,(32,32)),(16,(224,224))):
    enc=ClipImageEncoder(random_vit_state_dict(patch,seed=0,layers=1),device='cuda',max_batch=512)
    x=torch.randint(0,256,(512,h,w,3),dtype=torch.uint8,device='cuda')
    from torch.profiler import profile, ProfilerActivity
    enc(x); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        enc(x); torch.cuda.synchronize()
    for e in prof.events():
        if 'patchify' in e.name or 'im2col' in e.name: print(patch,(h,w),e.name[:40], e.device_time_total if hasattr(e,'device_time_total') else e.cuda_time_total,'us')
PY
