"""Stand-alone diagnostics for the tcgen05 GEMM / LN / attention building blocks (run on the GPU box).
Each case runs in its own process under a timeout by tools/run_debug.sh so that a trap cannot mask later cases."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eoe_b200 import encoder as E, _lib as L

def gemm_case(M, N, K, dtype, epi=0):
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ W.float().t() + bias
    if epi == L.EOE_EPI_BIAS_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    if epi == L.EOE_EPI_BIAS_RESIDUAL_F32:
        out = torch.randn(M, N, device="cuda")
        ref = ref + out
        E.gemm(A, W, bias, epi, out=out)
        got = out
    else:
        got = E.gemm(A, W, bias, epi).float()
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"gemm M={M} N={N} K={K} {dtype} epi={epi}: max_abs_err={err:.4e} rel_l2={rel:.3e} ref_max={ref.abs().max().item():.3f}", flush=True)
    if rel > 1e-2:
        bad = ((got - ref).abs() > 0.1 * ref.abs().max()).nonzero()
        print("  first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], flush=True)
        print("  got[0,:8]", got[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist(), flush=True)

def bench_gemm(M, N, K, dtype, epi=0, iters=20):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=torch.float32 if epi == 2 else dtype, device="cuda")
    for _ in range(3): E.gemm(A, W, bias, epi, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): E.gemm(A, W, bias, epi, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    t0 = time.time()
    for _ in range(3): torch.matmul(A, W.t())
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters): torch.matmul(A, W.t())
    b.record(); torch.cuda.synchronize()
    ms_t = a.elapsed_time(b) / iters
    fl = 2.0 * M * N * K
    print(f"bench M={M} N={N} K={K} epi={epi}: ours {ms:.4f} ms = {fl/ms/1e9:.1f} TFLOP/s | cuBLAS {ms_t:.4f} ms = {fl/ms_t/1e9:.1f} TFLOP/s", flush=True)

def ln_case():
    x = torch.randn(1000, 768, device="cuda") * 2 + 0.3
    w = torch.randn(768, device="cuda"); b = torch.randn(768, device="cuda")
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-5)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        got = E.layernorm(x, w, b, dt).float()
        print("layernorm", dt, (got - ref).abs().max().item(), flush=True)

def attn_case(B, Lseq, dtype):
    torch.manual_seed(1)
    heads, W = 12, 768
    qkv = (torch.randn(B * Lseq, 3 * W, device="cuda")).to(dtype)
    got = E.attention(qkv, B, Lseq, heads).float()
    q, k, v = qkv.float().split(W, dim=-1)
    q = q.reshape(B, Lseq, heads, 64).transpose(1, 2); k = k.reshape(B, Lseq, heads, 64).transpose(1, 2); v = v.reshape(B, Lseq, heads, 64).transpose(1, 2)
    ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(B * Lseq, W)
    torch.cuda.synchronize()
    print(f"attention B={B} L={Lseq} {dtype}: max_abs_err={(got-ref).abs().max().item():.4e} rel_l2={((got-ref).norm()/ref.norm()).item():.3e}", flush=True)

if __name__ == "__main__":
    case = sys.argv[1]
    bf, hf = torch.bfloat16, torch.float16
    if case == "gemm_small":
        gemm_case(128, 256, 64, bf); gemm_case(128, 256, 128, bf); gemm_case(128, 256, 768, bf)
        gemm_case(256, 512, 768, bf); gemm_case(100, 256, 768, bf); gemm_case(1000, 768, 3072, hf)
    elif case == "gemm_epi":
        for epi in (0, 1, 2):
            gemm_case(6400, 2304 if epi == 0 else (3072 if epi == 1 else 768), 768 if epi < 2 else 3072, bf, epi)
        gemm_case(25216, 2304, 768, hf, 0)
    elif case == "gemm_bench":
        for (M, N, K, epi) in [(50432, 2304, 768, 0), (50432, 768, 768, 2), (50432, 3072, 768, 1), (50432, 768, 3072, 2), (8192, 8192, 8192, 0)]:
            bench_gemm(M, N, K, bf, epi)
    elif case == "ln":
        ln_case()
    elif case == "attn_bench":
        for (B, Lseq) in [(512, 197), (512, 50)]:
            qkv = torch.randn(B * Lseq, 2304, device="cuda").to(bf)
            for _ in range(3): E.attention(qkv, B, Lseq, 12)
            torch.cuda.synchronize()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): E.attention(qkv, B, Lseq, 12)
            b2.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b2) / 20
            fl = 4.0 * B * 12 * Lseq * Lseq * 64
            print(f"attention bench B={B} L={Lseq}: {ms*1e3:.1f} us, {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    elif case == "attn":
        attn_case(2, 50, bf); attn_case(3, 197, bf); attn_case(2, 197, hf); attn_case(1, 64, bf); attn_case(1, 208, bf); attn_case(2, 17, hf)
