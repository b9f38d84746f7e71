"""GPU: no entry point writes outside the extents include/eoe_b200.h documents.

compute-sanitizer is closed on this pool (profiles/r2_sanitizer_unavailable.md), so this is the repo's own bounds check:
every output and workspace of a call is carved out of a larger allocation whose guard bands (4 KiB before and after) carry
a sentinel byte pattern; after the call (ragged sizes: rows that fill no tile, n = 1, partially filled last tiles) every
guard byte must be untouched and every output fully written where the header says so.  Calls go through the raw C ABI
(ctypes) because the Python wrappers allocate their own outputs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GUARD = 4096
SENT = 0xA5


class Guarded:
    """A tensor view of `nbytes` between two sentinel bands."""

    def __init__(self, shape, dtype, fill=None):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.pad = (-n) % 256
        self.raw = torch.full((GUARD + n + self.pad + GUARD,), SENT, dtype=torch.uint8, device=DEV)
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(shape)
        if fill is not None:
            self.t.copy_(fill)

    def ptr(self):
        return C.c_void_p(self.t.data_ptr())

    def check(self, name):
        torch.cuda.synchronize()
        head, tail = self.raw[:GUARD], self.raw[self.raw.numel() - GUARD - self.pad:]
        assert bool((head == SENT).all()), f"{name}: bytes BEFORE the buffer were written"
        assert bool((tail == SENT).all()), f"{name}: bytes AFTER the buffer were written"


def _lib():
    from eoe_b200 import _lib as L
    return L, L.lib()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n,d", [(1, 256), (257, 256), (33, 100), (1000, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_hsc_and_bce_write_only_their_outputs(n, d, dtype):
    L, lib = _lib()
    z = (0.05 * torch.randn(n, d, device=DEV)).to(dtype)
    y = torch.randint(0, 2, (n,), device=DEV)
    loss, scores, grad = Guarded((1,), torch.float32), Guarded((n,), torch.float32), Guarded((n, d), dtype)
    ws = Guarded((L.EOE_HEAD_WS_BYTES,), torch.uint8, fill=0)
    L.check(lib.eoe_hsc_fwd_bwd(L.ptr(z), L.DTYPE_CODE[dtype], L.ptr(y), n, d, 0, loss.ptr(), scores.ptr(), grad.ptr(),
                                ws.ptr(), _stream()), "hsc")
    for g, nm in ((loss, "loss"), (scores, "scores"), (grad, "grad"), (ws, "head_ws")):
        g.check("eoe_hsc_fwd_bwd " + nm)
    assert torch.isfinite(scores.t).all() and torch.isfinite(grad.t.float()).all()
    assert bool((ws.t[:4] == 0).all())                                 # the ticket is handed back re-armed (partials are scratch)
    x = torch.randn(n, 1, device=DEV).to(dtype)
    gx = Guarded((n, 1), dtype)
    L.check(lib.eoe_bce_fwd_bwd(L.ptr(x), L.DTYPE_CODE[dtype], L.ptr(y), n, 0, loss.ptr(), scores.ptr(), gx.ptr(), ws.ptr(),
                                _stream()), "bce")
    for g, nm in ((loss, "loss"), (scores, "scores"), (gx, "grad"), (ws, "head_ws")):
        g.check("eoe_bce_fwd_bwd " + nm)


@pytest.mark.parametrize("n,K", [(77, 10), (2050, 30), (2049, 2), (130, 100)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_clip_heads_write_only_their_outputs(n, K, dtype):
    L, lib = _lib()
    d = 512
    z = torch.randn(n, d, device=DEV).to(dtype)
    c = torch.nn.functional.normalize(torch.randn(K, d, device=DEV), dim=-1)
    y = torch.randint(0, 2, (n,), device=DEV)
    scores, loss, grad = Guarded((n,), torch.float32), Guarded((1,), torch.float32), Guarded((n, d), dtype)
    ws = Guarded((L.EOE_HEAD_WS_BYTES,), torch.uint8, fill=0)
    L.check(lib.eoe_clip_score(L.ptr(z), L.DTYPE_CODE[dtype], L.ptr(c), n, d, K, 100.0, scores.ptr(), _stream()), "clip_score")
    scores.check("eoe_clip_score scores")
    assert torch.isfinite(scores.t).all()
    for loo in (0, 1):
        L.check(lib.eoe_clip_oe_loss_fwd_bwd(L.ptr(z), L.DTYPE_CODE[dtype], L.ptr(c), L.ptr(y), n, d, K, 100.0, 0, loo,
                                             loss.ptr(), grad.ptr(), ws.ptr(), _stream()), "clip_oe")
        for g, nm in ((loss, "loss"), (grad, "grad"), (ws, "head_ws")):
            g.check("eoe_clip_oe_loss_fwd_bwd " + nm)
        assert torch.isfinite(grad.t.float()).all()


@pytest.mark.parametrize("n", [1, 2, 3001, 16384, 16385, 70001, 655360, 655361])      # 655 360: last size with 2048-key sort tiles
def test_auc_writes_only_its_outputs_and_workspace(n):
    L, lib = _lib()
    rng = np.random.default_rng(n)
    s = torch.from_numpy((1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)).to(DEV)
    y = torch.from_numpy((rng.random(n) < 0.4).astype(np.int64)).to(DEV)
    nbytes = lib.eoe_auc_workspace_bytes(n)
    ws = Guarded((nbytes,), torch.uint8)
    out, info = Guarded((2,), torch.float64), Guarded((8,), torch.int64)
    fpr, tpr = Guarded((n + 1,), torch.float64), Guarded((n + 1,), torch.float64)
    thr, pthr = Guarded((n + 1,), torch.float32), Guarded((n,), torch.float32)
    prec, rec = Guarded((n + 1,), torch.float64), Guarded((n + 1,), torch.float64)
    L.check(lib.eoe_auc(L.ptr(s), L.EOE_F32, L.ptr(y), n, L.EOE_AUC_WITH_PRC, ws.ptr(), nbytes, out.ptr(), info.ptr(), fpr.ptr(),
                        tpr.ptr(), thr.ptr(), prec.ptr(), rec.ptr(), pthr.ptr(), _stream()), "auc")
    for g, nm in ((ws, "workspace"), (out, "auc_out"), (info, "info"), (fpr, "fpr"), (tpr, "tpr"), (thr, "thr"),
                  (pthr, "prc_thr"), (prec, "prec"), (rec, "rec")):
        g.check(f"eoe_auc(n={n}) " + nm)


@pytest.mark.parametrize("M,N,K", [(1, 256, 64), (257, 768, 768), (100, 2304, 768), (515, 768, 3072)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_and_attention_write_only_their_outputs(M, N, K, dtype):
    L, lib = _lib()
    A = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(dtype)
    bias = torch.randn(N, device=DEV)
    out16 = Guarded((M, N), dtype)
    for epi in (L.EOE_EPI_BIAS, L.EOE_EPI_BIAS_QUICKGELU):
        L.check(lib.eoe_gemm(L.ptr(A), L.ptr(W), L.ptr(bias), out16.ptr(), M, N, K, L.DTYPE_CODE[dtype], epi, None, 0, _stream()), "gemm")
        out16.check(f"eoe_gemm epi {epi}")
    out32 = Guarded((M, N), torch.float32, fill=torch.zeros(M, N, device=DEV))
    L.check(lib.eoe_gemm(L.ptr(A), L.ptr(W), L.ptr(bias), out32.ptr(), M, N, K, L.DTYPE_CODE[dtype], L.EOE_EPI_BIAS_RESIDUAL_F32,
                         None, 0, _stream()), "gemm residual")
    out32.check("eoe_gemm residual")
    ref = A.float() @ W.float().t() + bias
    assert ((out32.t - ref).norm() / ref.norm()).item() < 1e-4
    if N % 128 == 0 and N <= 1024:
        # residual + statistics epilogue: x in place, 16-bit centred copy, chunk sums, shifts
        x = Guarded((M, N), torch.float32, fill=torch.randn(M, N, device=DEV))
        xb, stats, shift = Guarded((M, N), dtype), Guarded((M, N // 128, 2), torch.float32), Guarded((M,), torch.float32)
        L.check(lib.eoe_gemm_residual_stats(L.ptr(A), L.ptr(W), L.ptr(bias), None, x.ptr(), xb.ptr(), stats.ptr(), shift.ptr(),
                                            M, N, K, L.DTYPE_CODE[dtype], _stream()), "gemm_residual_stats")
        for g, nm in ((x, "x"), (xb, "xb"), (stats, "stats"), (shift, "shift")):
            g.check("eoe_gemm_residual_stats " + nm)


@pytest.mark.parametrize("B,Lseq", [(1, 197), (3, 197), (2, 50), (1, 64), (2, 17), (1, 208)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_writes_only_its_output(B, Lseq, dtype):
    L, lib = _lib()
    qkv = torch.randn(B * Lseq, 3 * 768, device=DEV).to(dtype)
    out = Guarded((B * Lseq, 768), dtype)
    L.check(lib.eoe_attention(L.ptr(qkv), out.ptr(), B, Lseq, 12, L.DTYPE_CODE[dtype], _stream()), "attention")
    out.check("eoe_attention out")
    assert torch.isfinite(out.t.float()).all()


@pytest.mark.parametrize("patch,B", [(32, 3), (16, 2)])
def test_encoder_stays_inside_its_workspace(patch, B):
    """eoe_vit_encode with a workspace of EXACTLY eoe_vit_workspace_bytes between guard bands, max_batch == B (every buffer at
    its tightest), fp32 and uint8 inputs, features + fused scores."""
    from eoe_b200.encoder import ClipImageEncoder
    from oracle import vit as ovit
    L, lib = _lib()
    sd = ovit.synth_state_dict(patch, seed=2, layers=2)
    enc = ClipImageEncoder(sd, device=DEV, max_batch=B)
    nbytes = lib.eoe_vit_workspace_bytes(C.byref(enc._w), B)
    ws = Guarded((nbytes + 1024,), torch.uint8)
    base = ws.t.data_ptr() + ((-ws.t.data_ptr()) % 1024)
    plan = C.c_void_p()
    L.check(lib.eoe_vit_plan_create(C.byref(enc._w), B, C.c_void_p(base), nbytes, C.byref(plan)), "plan")
    try:
        imgs = torch.randn(B, 3, 224, 224, device=DEV)
        text = torch.nn.functional.normalize(torch.randn(10, 512, device=DEV), dim=-1)
        feats, scores = Guarded((B, 512), torch.float32), Guarded((B,), torch.float32)
        L.check(lib.eoe_vit_encode(plan, L.ptr(imgs), B, feats.ptr(), L.ptr(text), 10, 100.0, scores.ptr(), _stream()), "encode")
        for g, nm in ((ws, "workspace"), (feats, "features"), (scores, "scores")):
            g.check("eoe_vit_encode " + nm)
        assert torch.equal(feats.t, enc(imgs))                 # and the guarded run computes what the module computes
        u8 = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=DEV)
        L.check(lib.eoe_vit_encode_u8(plan, L.ptr(u8), L.EOE_LAYOUT_NHWC, enc._mean, enc._std, B, feats.ptr(), L.ptr(text), 10,
                                      100.0, scores.ptr(), _stream()), "encode_u8")
        for g, nm in ((ws, "workspace"), (feats, "features"), (scores, "scores")):
            g.check("eoe_vit_encode_u8 " + nm)
    finally:
        torch.cuda.synchronize()
        lib.eoe_vit_plan_destroy(plan)


# ------------------------------------------------------------------------------------------ round 2: new paths
@pytest.mark.parametrize("n,K", [(16384, 30), (16385, 10), (20001, 32), (16500, 2)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_clip_tcgen05_heads_write_only_their_outputs(n, K, dtype):
    """16-bit rows at n >= 16 384 take the tcgen05 kernels (TMA tile stores of dz, rows past n clipped by the tensor map)."""
    L, lib = _lib()
    d = 512
    z = torch.randn(n, d, device=DEV).to(dtype)
    c = torch.nn.functional.normalize(torch.randn(K, d, device=DEV), dim=-1)
    y = torch.randint(0, 2, (n,), device=DEV)
    scores, loss, grad = Guarded((n,), torch.float32), Guarded((1,), torch.float32), Guarded((n, d), dtype)
    ws = Guarded((L.EOE_HEAD_WS_BYTES,), torch.uint8, fill=0)
    L.check(lib.eoe_clip_score(L.ptr(z), L.DTYPE_CODE[dtype], L.ptr(c), n, d, K, 100.0, scores.ptr(), _stream()), "clip_score")
    scores.check("eoe_clip_score (tcgen05) scores")
    assert torch.isfinite(scores.t).all()
    for loo in (0, 1):
        L.check(lib.eoe_clip_oe_loss_fwd_bwd(L.ptr(z), L.DTYPE_CODE[dtype], L.ptr(c), L.ptr(y), n, d, K, 100.0, 0, loo,
                                             loss.ptr(), grad.ptr(), ws.ptr(), _stream()), "clip_oe")
        for g, nm in ((loss, "loss"), (grad, "grad"), (ws, "head_ws")):
            g.check("eoe_clip_oe_loss_fwd_bwd (tcgen05) " + nm)
        assert torch.isfinite(grad.t.float()).all() and bool((ws.t[:4] == 0).all())


@pytest.mark.parametrize("M,N,K", [(1, 256, 64), (257, 768, 768), (100, 2304, 768), (515, 768, 3072)])
def test_split_gemm_and_attention_write_only_their_outputs(M, N, K):
    """operand dtype EOE_F16X2: operands [rows, 2K], 16-bit outputs [M, 2N] (hi | lo tiles, two TMA stores per 64 columns)."""
    from eoe_b200 import encoder as E
    L, lib = _lib()
    A = E.split_f16(torch.randn(M, K, device=DEV) * 0.5)
    W = E.split_f16(torch.randn(N, K, device=DEV) * 0.05)
    bias = torch.randn(N, device=DEV)
    out16 = Guarded((M, 2 * N), torch.float16)
    for epi in (L.EOE_EPI_BIAS, L.EOE_EPI_BIAS_QUICKGELU):
        L.check(lib.eoe_gemm(L.ptr(A), L.ptr(W), L.ptr(bias), out16.ptr(), M, N, K, L.EOE_F16X2, epi, None, 0, _stream()), "gemm")
        out16.check(f"eoe_gemm split epi {epi}")
        assert torch.isfinite(out16.t.float()).all()
    if N % 128 == 0 and N <= 1024:
        x = Guarded((M, N), torch.float32, fill=torch.randn(M, N, device=DEV))
        xb, stats, shift = Guarded((M, 2 * N), torch.float16), Guarded((M, N // 128, 2), torch.float32), Guarded((M,), torch.float32)
        L.check(lib.eoe_gemm_residual_stats(L.ptr(A), L.ptr(W), L.ptr(bias), None, x.ptr(), xb.ptr(), stats.ptr(), shift.ptr(),
                                            M, N, K, L.EOE_F16X2, _stream()), "gemm_residual_stats")
        for g, nm in ((x, "x"), (xb, "xb"), (stats, "stats"), (shift, "shift")):
            g.check("eoe_gemm_residual_stats split " + nm)
    for B, Lseq in ((2, 197), (3, 50), (1, 17)):
        qkv = E.split_f16(torch.randn(B * Lseq, 3 * 768, device=DEV))
        out = Guarded((B * Lseq, 2 * 768), torch.float16)
        L.check(lib.eoe_attention(L.ptr(qkv), out.ptr(), B, Lseq, 12, L.EOE_F16X2, _stream()), "attention")
        out.check("eoe_attention split out")
        assert torch.isfinite(out.t.float()).all()
        if Lseq <= 100:
            outc = Guarded((B * Lseq, 2 * 768), torch.float16)
            L.check(lib.eoe_attention_causal(L.ptr(qkv), outc.ptr(), B, Lseq, 12, L.EOE_F16X2, _stream()), "attention_causal")
            outc.check("eoe_attention_causal split out")


@pytest.mark.parametrize("patch,B", [(32, 3), (16, 2)])
def test_encoder_precise_mode_stays_inside_its_workspace(patch, B):
    """the precise mode doubles every 16-bit buffer: eoe_vit_workspace_bytes must account for all of them"""
    from eoe_b200.encoder import ClipImageEncoder
    from oracle import vit as ovit
    L, lib = _lib()
    sd = ovit.synth_state_dict(patch, seed=2, layers=2)
    enc = ClipImageEncoder(sd, device=DEV, max_batch=B, operand_dtype="f16x2")
    nbytes = lib.eoe_vit_workspace_bytes(C.byref(enc._w), B)
    ws = Guarded((nbytes + 1024,), torch.uint8)
    base = ws.t.data_ptr() + ((-ws.t.data_ptr()) % 1024)
    plan = C.c_void_p()
    L.check(lib.eoe_vit_plan_create(C.byref(enc._w), B, C.c_void_p(base), nbytes, C.byref(plan)), "plan")
    try:
        imgs = torch.randn(B, 3, 224, 224, device=DEV)
        text = torch.nn.functional.normalize(torch.randn(10, 512, device=DEV), dim=-1)
        feats, scores = Guarded((B, 512), torch.float32), Guarded((B,), torch.float32)
        L.check(lib.eoe_vit_encode(plan, L.ptr(imgs), B, feats.ptr(), L.ptr(text), 10, 100.0, scores.ptr(), _stream()), "encode")
        for g, nm in ((ws, "workspace"), (feats, "features"), (scores, "scores")):
            g.check("eoe_vit_encode (f16x2) " + nm)
        assert torch.equal(feats.t, enc(imgs))
        u8 = torch.randint(0, 256, (B, 64, 100, 3), dtype=torch.uint8, device=DEV)
        L.check(lib.eoe_vit_encode_u8_resize(plan, L.ptr(u8), 64, 100, enc._mean, enc._std, B, feats.ptr(), L.ptr(text), 10,
                                             100.0, scores.ptr(), _stream()), "encode_u8_resize")
        for g, nm in ((ws, "workspace"), (feats, "features"), (scores, "scores")):
            g.check("eoe_vit_encode_u8_resize (f16x2) " + nm)
    finally:
        torch.cuda.synchronize()
        lib.eoe_vit_plan_destroy(plan)
