#!/usr/bin/env python
"""Stand-alone timing of the encoder's GEMM shapes with the kernel's diagnostic flags (eoe_debug_set):
0 normal, 1 main loop only (epilogue just releases the accumulators), 2 full epilogue without global stores,
16 two CTA pairs per cluster sharing W by TMA multicast, bits 8.. grid size in CTA pairs.  Answers: is a shape main-loop-, epilogue- or store-bound?"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eoe_b200 import _lib as L, encoder as E  # noqa: E402


def main():
    lib = L.lib()
    lib.eoe_debug_set.argtypes = [C.c_int]
    lib.eoe_debug_set.restype = None
    M = 512 * 197
    shapes = [("qkv", 2304, 768, 0), ("c_fc", 3072, 768, 1), ("out_proj", 768, 768, 2), ("c_proj", 768, 3072, 2)]
    flags = [int(f) for f in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2".split(","))]
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, N, K, epi in shapes:
        # two alternating A / out buffers so that successive launches do not find their operands in L2
        As = [(torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(2)]
        W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        outs = [torch.zeros(M, N, dtype=torch.float32 if epi == 2 else torch.bfloat16, device="cuda") for _ in range(2)]
        res[name] = {}
        for f in flags:
            lib.eoe_debug_set(f)
            for i in range(4):
                E.gemm(As[i & 1], W, bias, epi, out=outs[i & 1])
            torch.cuda.synchronize()
            e0.record()
            n = 20
            for i in range(n):
                E.gemm(As[i & 1], W, bias, epi, out=outs[i & 1])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            res[name][f] = {"us": round(us, 1), "tflops": round(2.0 * M * N * K / us / 1e6)}
            if f & 4:
                buf = (C.c_ulonglong * 4)()
                lib.eoe_debug_gemm_prof(buf)
                t = max(1, buf[2])
                res[name][f]["epilogue_clks_per_tile"] = {"wait_for_accumulator": buf[0] // t, "epilogue": buf[1] // t, "tiles": buf[2],
                                                          "mma_floor_clks_per_tile": 128 * 4 * (K // 64)}
            if not isinstance(res[name], dict):
                continue
            pairs = (f >> 8) or 74
            res[name][f]["tflops_per_pair"] = round(2.0 * M * N * K / us / 1e6 / pairs, 2)
        lib.eoe_debug_set(0)
        del As, outs
    res["max_active_clusters"] = {"pairs_per_cluster_1": lib.eoe_debug_max_clusters(1), "pairs_per_cluster_2": lib.eoe_debug_max_clusters(2)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
