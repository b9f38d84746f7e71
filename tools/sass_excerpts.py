#!/usr/bin/env python
"""Per-kernel SASS evidence: for the tcgen05 kernels of libeoe_b200.so, the census of the mnemonics
that prove tcgen05 / TMEM / TMA use (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tile
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and a listing excerpt around the first occurrence of each.

    python tools/sass_excerpts.py > profiles/r2_sass_excerpts.md        (cuobjdump -sass; runs without a GPU)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eoe_b200", "libeoe_b200.so")
WANT = [  # (mangled-name regex, label): fp16 instantiations (the default operand dtype), 1 CTA pair per cluster
    (r"gemm_kernelILi4ELb0ELi1ELb0E", "gemm_kernel<EOE_EPI_LNFOLD_BIAS, fp16> (QKV with ln_1 folded in)"),
    (r"gemm_kernelILi8ELb0ELi1ELb0E", "gemm_kernel<EOE_EPI_LNFOLD_QUICKGELU_X1702, fp16> (c_fc with ln_2 folded in)"),
    (r"gemm_kernelILi7ELb0ELi1ELb0E", "gemm_kernel<EPI_RESIDUAL_STATS_ASYNC, fp16> (out-proj, residual + statistics)"),
    (r"gemm_kernelILi6ELb0ELi1ELb0E", "gemm_kernel<EOE_EPI_RESIDUAL_STATS, fp16> (c_proj, residual + statistics)"),
    (r"gemm_kernelILi3ELb0ELi1ELb0E", "gemm_kernel<EOE_EPI_PATCH_EMBED, fp16> (patch embedding)"),
    (r"attention_tc_kernelILb0ELi197ELb0E", "attention_tc_kernel<fp16, L = 197> (ViT-B/16)"),
    (r"attention_tc64_kernelILb0ELb0E", "attention_tc64_kernel<fp16> (L <= 64, ViT-B/32)"),
    # precise mode (operand dtype EOE_F16X2: split fp16 pairs, SPLIT = true)
    (r"gemm_kernelILi8ELb0ELi1ELb1E", "gemm_kernel<EOE_EPI_LNFOLD_QUICKGELU_X1702, SPLIT> (c_fc, precise mode: hi / lo output tiles)"),
    (r"gemm_kernelILi6ELb0ELi1ELb1E", "gemm_kernel<EOE_EPI_RESIDUAL_STATS, SPLIT> (c_proj, precise mode)"),
    (r"attention_tc_kernelILb0ELi197ELb1E", "attention_tc_kernel<fp16, L = 197, SPLIT> (precise mode: 3 S products, P as a pair, 3 P.V products)"),
    (r"attention_tc64_kernelILb0ELb1E", "attention_tc64_kernel<fp16, SPLIT> (precise mode, L <= 64)"),
    # CLIP heads for 16-bit rows (csrc/clip_head_sm100.cuh)
    (r"clip_score_tc_kernelILb1E", "cliptc::clip_score_tc_kernel<bf16> (zero-shot score head, 16-bit rows)"),
    (r"clip_oe_loss_tc_kernelILb1E", "cliptc::clip_oe_loss_tc_kernel<bf16> (OE loss + backward, 16-bit rows: G through TMEM, MN-major text tiles)"),
]
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH",
             "REDG", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
    starts = [(i, l.split("Function :")[1].strip()) for i, l in enumerate(sass) if "Function :" in l]
    from eoe_b200 import build
    print(f"# SASS excerpts of eoe_b200/libeoe_b200.so (source id {build.source_id()}, `cuobjdump -sass`, sm_100a)\n")
    print("Mnemonics: `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` / `STTM` = tcgen05.ld / tcgen05.st (TMEM), `UTMALDG` / `UTMASTG` = "
          "cp.async.bulk.tensor load / store (TMA), `UTCBAR` = tcgen05.commit -> mbarrier, `SYNCS` = mbarrier ops, "
          "`HMMA` = legacy mma.sync (absent from these kernels).\n")
    for rx, label in WANT:
        hit = [(i, n) for i, n in starts if re.search(rx, n)]
        if not hit:
            print(f"## {label}\n\nNOT FOUND ({rx})\n")
            continue
        i0, name = hit[0]
        i1 = min([j for j, _ in starts if j > i0] + [len(sass)])
        body = [l for l in sass[i0:i1] if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ins = [re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", l)).strip(" ;") for l in body]
        print(f"## {label}\n\n`{name}` — {len(ins)} instructions\n")
        print("| mnemonic | count |\n|---|---:|")
        for m in MNEMONICS:
            c = sum(1 for x in ins if re.search(r"(^|\s)" + re.escape(m), x))
            print(f"| {m} | {c} |")
        print()
        shown = set()
        for m in ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMASTG"):
            for k, x in enumerate(ins):
                if re.search(r"(^|\s)" + re.escape(m), x) and not any(abs(k - s) < 4 for s in shown):
                    shown.add(k)
                    print(f"first `{m}` (instruction {k}):\n\n```")
                    for y in ins[max(0, k - 2):k + 3]:
                        print("    " + y)
                    print("```\n")
                    break


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
