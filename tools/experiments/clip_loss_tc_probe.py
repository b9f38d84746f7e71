"""Where the tcgen05 CLIP loss kernel's time goes: eoe_debug_set bits 20 (no dz stores) / 21 (no dz arithmetic)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eoe_b200 import _lib, ops
n, d, K = 1 << 20, 512, 30
z = torch.randn(n, d, device="cuda").to(torch.bfloat16)
y = torch.randint(0, 2, (n,), device="cuda")
c = torch.nn.functional.normalize(torch.randn(K, d, device="cuda"), dim=-1)
for flags in (0, 1 << 20):
    _lib.lib().eoe_debug_set(flags)
    for _ in range(3): ops.clip_oe_fused(z, y, c, 0, False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.clip_oe_fused(z, y, c, 0, False)
    b.record(); torch.cuda.synchronize()
    print(f"flags {flags >> 20}: {a.elapsed_time(b) / 10:.4f} ms")
_lib.lib().eoe_debug_set(0)
