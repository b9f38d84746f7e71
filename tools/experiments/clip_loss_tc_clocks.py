"""Phase clocks of the tcgen05 CLIP loss kernel (CTA 0), from an INSTRUMENTED VARIANT of the library
(tools/_variants/libeoe_b200_prof.so, built from a copy of csrc/ with clock64 probes; the product build is untouched):
    EOE_B200_LIB=tools/_variants/libeoe_b200_prof.so python tools/experiments/clip_loss_tc_clocks.py"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eoe_b200 import _lib, ops
lib = _lib.lib()
n, d, K = 1 << 20, 512, 30
z = torch.randn(n, d, device="cuda").to(torch.bfloat16)
y = torch.randint(0, 2, (n,), device="cuda")
c = torch.nn.functional.normalize(torch.randn(K, d, device="cuda"), dim=-1)
for _ in range(3): ops.clip_oe_fused(z, y, c, 0, False)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
lib.eoe_debug_clip_prof.argtypes = [C.c_void_p, C.c_int]
lib.eoe_debug_clip_prof(buf, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ops.clip_oe_fused(z, y, c, 0, False); b.record(); torch.cuda.synchronize()
lib.eoe_debug_clip_prof(buf, 0)
v = list(buf)
m = max(v[8], 1)
out = {"ms": a.elapsed_time(b), "tiles_of_cta0": v[8],
       "B0_warp4_clk_per_tile": dict(zip(["wait_gready", "wait_bfull", "tmem_ld", "bempty+wait_store_read", "wait_full", "math", "fence+store_tail", "total"], [x / m for x in v[0:8]])),
       "B1_warp8_clk_per_tile": dict(zip(["wait_gready", "wait_bfull", "tmem_ld", "bempty+wait_store_read", "wait_full", "math", "fence+store_tail", "total"], [x / m for x in v[10:18]])),
       "F_warp0_clk_per_tile": dict(zip(["wait_full(norm pass)", "wait_tfull", "softmax..G", "wait_gfree", "total"], [x / m for x in v[20:25]])),
       "issuer_clk_per_tile": dict(zip(["wait_tempty", "wait_full(fwd)", "wait_gready", "wait_bempty", "total"], [x / m for x in v[30:35]])),
       "producer_wait_empty_clk_per_tile": v[40] / m}
print(json.dumps(out, indent=1))
