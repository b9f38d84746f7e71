"""CLIP ViT-B image encoder as a drop-in `nn.Module`: `model(imgs) -> features [B, 512]`, the contract the
reference trainer relies on (`image_features = model(imgs)`, src/eoe/training/ad_trainer.py:429,507, with
`model.forward = model.encode_image`, src/eoe/training/clip.py:33).

Weights are loaded from the reference's own state_dict keys (`visual.conv1.weight`,
`visual.transformer.resblocks.{i}.attn.in_proj_weight`, ...; clip_official/clip/model.py:395-402), either with
or without the `visual.` prefix.  The forward pass is one C call (`eoe_vit_encode`) that launches the hand-written
sm_100a kernels; torch only owns the buffers.  Inference only (the zero-shot path of configs 2/3).
"""
import ctypes as C
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib as L


def _strip(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if any(k.startswith("visual.") for k in sd):
        return {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    return dict(sd)


class ClipImageEncoder(nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", operand_dtype=torch.float16,
                 max_batch: int = 256, heads: Optional[int] = None, resolution: Optional[int] = None,
                 fold_layernorm: bool = True, input_mean=(0.48145466, 0.4578275, 0.40821073),
                 input_std=(0.26862954, 0.26130258, 0.27577711)):
        """fold_layernorm: fold ln_1 / ln_2 into the QKV / c_fc GEMMs (eoe_vit_fold_layernorm; DESIGN.md "LayerNorm
        fold") instead of launching stand-alone LayerNorm kernels.  Same math (fp32 statistics of the fp32 residual
        stream); the 16-bit rounding point moves from LN(x) to x and to W*ln_w.

        operand_dtype: torch.float16 (default; the reference's own GPU dtype), torch.bfloat16, or "f16x2" -- the PRECISE
        mode (EOE_F16X2): every stored 16-bit tensor is an fp16 (hi, lo) pair and products are hi*hi + lo*hi + hi*lo in
        fp32 accumulators: 3x the tensor work, and end-to-end scores within 1e-3 relative of the reference's fp32 scores
        on every image (tests/test_gpu_encoder.py::test_end_to_end_scores)."""
        super().__init__()
        self.split = operand_dtype in (L.F16X2, "split")
        if self.split:
            operand_dtype = torch.float16
        if operand_dtype not in (torch.bfloat16, torch.float16):
            raise L.EoeError('operand_dtype must be torch.bfloat16, torch.float16 or "f16x2"')
        sd = _strip(state_dict)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.EoeError("ClipImageEncoder needs a CUDA device (no CPU fallback)")
        conv = sd["conv1.weight"]
        self.width, _, self.patch, _ = conv.shape
        n_pos = sd["positional_embedding"].shape[0]
        grid = int(round((n_pos - 1) ** 0.5))
        self.resolution = resolution or grid * self.patch
        self.heads = heads or self.width // 64
        self.embed_dim = sd["proj"].shape[1]
        self.n_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
        self.operand_dtype = operand_dtype           # storage dtype of the 16-bit tensors (fp16 in the precise mode)
        self.operand_mode = L.F16X2 if self.split else {torch.float16: "f16", torch.bfloat16: "bf16"}[operand_dtype]
        self.max_batch = int(max_batch)
        self.device_ = dev
        self.fold_layernorm = bool(fold_layernorm) and self.width <= 768      # eoe_gemm_lnfold: K <= 768
        if self.split and not self.fold_layernorm:
            raise L.EoeError('operand_dtype "f16x2" needs the LayerNorm-folded path (fold_layernorm=True, width <= 768)')
        self._code = L.EOE_F16X2 if self.split else L.DTYPE_CODE[operand_dtype]
        # uint8 inputs get ToTensor + Normalize fused into the patchify kernel; defaults are CLIP's constants
        # (clip_official/clip/clip.py:64)
        self._mean = (C.c_float * 3)(*[float(v) for v in input_mean])
        self._std = (C.c_float * 3)(*[float(v) for v in input_std])

        def f32(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def op(t):
            t32 = t.detach().to(device=dev, dtype=torch.float32)
            if self.split:                        # [N, 2K] = [hi | lo], lo = rn(x - hi)
                hi = t32.to(torch.float16)
                return torch.cat([hi, (t32 - hi.float()).to(torch.float16)], dim=1).contiguous()
            return t32.to(operand_dtype).contiguous()

        # device-resident parameters in kernel layout (buffers so .to()/state_dict do not disturb pointers)
        self._keep = []
        w = L.VitWeights()
        w.patch, w.resolution, w.width, w.heads = self.patch, self.resolution, self.width, self.heads
        w.n_layers, w.embed_dim, w.operand_dtype = self.n_layers, self.embed_dim, self._code

        def put(t):
            self._keep.append(t)
            return t.data_ptr()

        w.conv1_w = put(op(conv.reshape(self.width, -1)))
        w.class_embedding = put(f32(sd["class_embedding"]))
        w.positional_embedding = put(f32(sd["positional_embedding"]))
        w.ln_pre_w, w.ln_pre_b = put(f32(sd["ln_pre.weight"])), put(f32(sd["ln_pre.bias"]))
        w.ln_post_w, w.ln_post_b = put(f32(sd["ln_post.weight"])), put(f32(sd["ln_post.bias"]))
        w.proj = put(f32(sd["proj"]))
        layers = (L.VitLayer * self.n_layers)()
        for i in range(self.n_layers):
            p = f"transformer.resblocks.{i}."
            l = layers[i]
            l.ln_1_w, l.ln_1_b = put(f32(sd[p + "ln_1.weight"])), put(f32(sd[p + "ln_1.bias"]))
            l.in_proj_w, l.in_proj_b = put(op(sd[p + "attn.in_proj_weight"])), put(f32(sd[p + "attn.in_proj_bias"]))
            l.out_proj_w, l.out_proj_b = put(op(sd[p + "attn.out_proj.weight"])), put(f32(sd[p + "attn.out_proj.bias"]))
            l.ln_2_w, l.ln_2_b = put(f32(sd[p + "ln_2.weight"])), put(f32(sd[p + "ln_2.bias"]))
            l.c_fc_w, l.c_fc_b = put(op(sd[p + "mlp.c_fc.weight"])), put(f32(sd[p + "mlp.c_fc.bias"]))
            l.c_proj_w, l.c_proj_b = put(op(sd[p + "mlp.c_proj.weight"])), put(f32(sd[p + "mlp.c_proj.bias"]))
            if self.fold_layernorm:
                l.in_proj_wf, l.in_proj_c1, l.in_proj_c2 = self._fold(
                    f32(sd[p + "attn.in_proj_weight"]), l.ln_1_w, l.ln_1_b, l.in_proj_b, put)
                l.c_fc_wf, l.c_fc_c1, l.c_fc_c2 = self._fold(
                    f32(sd[p + "mlp.c_fc.weight"]), l.ln_2_w, l.ln_2_b, l.c_fc_b, put)
                if i < self.n_layers - 1:          # the c_fc epilogue then emits 1.702 * QuickGELU (2 multiplies fewer per element)
                    l.c_proj_w_div1702 = put(op(f32(sd[p + "mlp.c_proj.weight"]) / L.GELU_SLOPE))
        w.layers_host = C.cast(layers, C.POINTER(L.VitLayer))
        self._layers, self._w = layers, w

        lib = L.lib()
        nbytes = lib.eoe_vit_workspace_bytes(C.byref(w), self.max_batch)
        if nbytes == 0:
            raise L.EoeError("unsupported ViT configuration for eoe_vit_* (see include/eoe_b200.h)")
        self._ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        off = (-self._ws.data_ptr()) % 1024
        self._ws_ptr = self._ws.data_ptr() + off
        plan = C.c_void_p()
        L.check(lib.eoe_vit_plan_create(C.byref(w), self.max_batch, C.c_void_p(self._ws_ptr), nbytes, C.byref(plan)),
                "eoe_vit_plan_create")
        self._plan = plan

    def _fold(self, w32, ln_w_ptr, ln_b_ptr, bias_ptr, put):
        """eoe_vit_fold_layernorm on the fp32 master weights -> device pointers (folded W, c1, c2)."""
        N, K = w32.shape
        wf = torch.empty(N, (2 if self.split else 1) * K, dtype=self.operand_dtype, device=w32.device)
        c1 = torch.empty(N, dtype=torch.float32, device=w32.device)
        c2 = torch.empty(N, dtype=torch.float32, device=w32.device)
        L.check(L.lib().eoe_vit_fold_layernorm(L.ptr(w32), C.c_void_p(ln_w_ptr), C.c_void_p(ln_b_ptr), C.c_void_p(bias_ptr),
                                               N, K, self._code, L.ptr(wf), L.ptr(c1), L.ptr(c2),
                                               L.stream_ptr(w32.device)), "eoe_vit_fold_layernorm")
        torch.cuda.current_stream(w32.device).synchronize()      # w32 is a temporary
        return put(wf), put(c1), put(c2)

    def __del__(self):
        try:
            plan = self.__dict__.get("_plan")
            if plan:
                self.__dict__["_plan"] = None       # not through nn.Module.__setattr__: it may be half torn down at exit
                L.lib().eoe_vit_plan_destroy(plan)
        except Exception:
            pass

    @property
    def flops_per_image(self) -> float:
        g2 = (self.resolution // self.patch) ** 2
        Ls = g2 + 1
        W = self.width
        per_layer = 2 * Ls * W * 12 * W + 4 * Ls * Ls * W
        return float(2 * g2 * 3 * self.patch ** 2 * W + self.n_layers * per_layer + 2 * W * self.embed_dim)

    def _prep(self, imgs):
        """-> (contiguous tensor, layout or None).  float tensors [B,3,R,R] are already normalised (what the reference
        hands to model(imgs), ad_trainer.py:507); uint8 tensors [B,3,R,R] or [B,R,R,3] are raw pixels at the model's
        resolution; uint8 [B,H,W,3] of any other size goes through CLIP's Resize(bicubic) + CenterCrop on the device."""
        L.require_cuda(imgs)
        R = self.resolution
        if imgs.dim() != 4:
            raise L.EoeError(f"images must be 4-d, got {tuple(imgs.shape)}")
        if imgs.dtype == torch.uint8:
            if tuple(imgs.shape[1:]) == (3, R, R):
                return imgs.detach().contiguous(), L.EOE_LAYOUT_NCHW
            if tuple(imgs.shape[1:]) == (R, R, 3):
                return imgs.detach().contiguous(), L.EOE_LAYOUT_NHWC
            if imgs.shape[3] == 3:              # raw decoded images of another size: Resize + CenterCrop run on the device too
                return imgs.detach().contiguous(), L.LAYOUT_RESIZE
            raise L.EoeError(f"uint8 images must be [B,3,{R},{R}], [B,{R},{R},3] or raw [B,H,W,3], got {tuple(imgs.shape)}")
        if tuple(imgs.shape[1:]) != (3, R, R):
            raise L.EoeError(f"images must be [B,3,{R},{R}], got {tuple(imgs.shape)}")
        return imgs.detach().to(torch.float32).contiguous(), None     # encode_image casts to the weight dtype (model.py:337)

    def _encode(self, imgs, layout, n, feats, text, K, scale, scores):
        lib = L.lib()
        if layout is None:
            L.check(lib.eoe_vit_encode(self._plan, L.ptr(imgs), n, L.ptr(feats), L.ptr(text), K, float(scale), L.ptr(scores),
                                       L.stream_ptr(imgs.device)), "eoe_vit_encode")
        elif layout == L.LAYOUT_RESIZE:
            L.check(lib.eoe_vit_encode_u8_resize(self._plan, L.ptr(imgs), imgs.shape[1], imgs.shape[2], self._mean, self._std, n,
                                                 L.ptr(feats), L.ptr(text), K, float(scale), L.ptr(scores),
                                                 L.stream_ptr(imgs.device)), "eoe_vit_encode_u8_resize")
        else:
            L.check(lib.eoe_vit_encode_u8(self._plan, L.ptr(imgs), layout, self._mean, self._std, n, L.ptr(feats), L.ptr(text),
                                          K, float(scale), L.ptr(scores), L.stream_ptr(imgs.device)), "eoe_vit_encode_u8")

    @torch.no_grad()
    def forward(self, imgs: torch.Tensor) -> torch.Tensor:
        imgs, layout = self._prep(imgs)
        B = imgs.shape[0]
        feats = torch.empty(B, self.embed_dim, dtype=torch.float32, device=imgs.device)
        for s in range(0, B, self.max_batch):
            n = min(self.max_batch, B - s)
            self._encode(imgs[s:s + n], layout, n, feats[s:s + n], None, 0, 100.0, None)
        return feats

    encode_image = forward

    GEMM_KINDS = ("patch_embed", "qkv", "out_proj", "c_fc", "c_proj")

    def profile(self, enable: bool):
        """Bracket every GEMM launch of subsequent forward calls with CUDA events (see eoe_vit_profile_enable)."""
        L.check(L.lib().eoe_vit_profile_enable(self._plan, int(enable)), "eoe_vit_profile_enable")

    def profile_read(self):
        """{kind: (ms, launches, flops)} accumulated since the last read; synchronises on the recorded events."""
        ms, n, fl = (C.c_double * 5)(), (C.c_int64 * 5)(), (C.c_double * 5)()
        L.check(L.lib().eoe_vit_profile_read(self._plan, ms, n, fl), "eoe_vit_profile_read")
        return {k: (ms[i], n[i], fl[i]) for i, k in enumerate(self.GEMM_KINDS)}

    @torch.no_grad()
    def score(self, imgs: torch.Tensor, center: torch.Tensor, scale: float = 100.0, out: Optional[torch.Tensor] = None
              ) -> torch.Tensor:
        """Fused zero-shot path: encoder + ADClipTrainer.compute_anomaly_score (clip.py:66-79) -> scores [B]."""
        imgs, layout = self._prep(imgs)
        B = imgs.shape[0]
        text = center.detach().to(device=imgs.device, dtype=torch.float32).contiguous()
        scores = out if out is not None else torch.empty(B, dtype=torch.float32, device=imgs.device)
        for s in range(0, B, self.max_batch):
            n = min(self.max_batch, B - s)
            self._encode(imgs[s:s + n], layout, n, None, text, text.shape[0], scale, scores[s:s + n])
        return scores


# thin functional wrappers over the exported building blocks (parity-tested one by one)
def split_f16(t):
    """[rows, C] fp32 -> [rows, 2C] fp16 = [hi | lo], the EOE_F16X2 storage of a matrix; join_f16 is its inverse (fp32)."""
    t32 = t.to(torch.float32)
    hi = t32.to(torch.float16)
    return torch.cat([hi, (t32 - hi.float()).to(torch.float16)], dim=1).contiguous()


def join_f16(t):
    C2 = t.shape[1] // 2
    return t[:, :C2].float() + t[:, C2:].float()


def gemm(A, W, bias=None, epilogue=L.EOE_EPI_BIAS, out=None, aux=None, aux_i=0, split=False):
    """split=True: A [M, 2K], W [N, 2K] and 16-bit outputs [M, 2N] are fp16 (hi | lo) pairs (EOE_F16X2)."""
    L.require_cuda(A, W)
    M, K = A.shape
    N = W.shape[0]
    if split:
        K //= 2
    if out is None:
        out = torch.empty(M, (2 if split else 1) * N, dtype=A.dtype, device=A.device)
    L.check(L.lib().eoe_gemm(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(out), M, N, K, L.EOE_F16X2 if split else L.DTYPE_CODE[A.dtype], epilogue,
                             L.ptr(aux), int(aux_i), L.stream_ptr(A.device)), "eoe_gemm")
    return out


def fold_layernorm(w32, ln_w, ln_b, bias, operand_dtype=torch.bfloat16):
    """(W*ln_w rounded to operand dtype, c1, c2) of eoe_vit_fold_layernorm; operand_dtype "f16x2": W*ln_w as [N, 2K] pairs."""
    L.require_cuda(w32, ln_w, ln_b)
    N, K = w32.shape
    split = operand_dtype == L.F16X2
    wf = torch.empty(N, (2 if split else 1) * K, dtype=torch.float16 if split else operand_dtype, device=w32.device)
    c1 = torch.empty(N, dtype=torch.float32, device=w32.device)
    c2 = torch.empty(N, dtype=torch.float32, device=w32.device)
    L.check(L.lib().eoe_vit_fold_layernorm(L.ptr(w32), L.ptr(ln_w), L.ptr(ln_b), L.ptr(bias), N, K,
                                           L.EOE_F16X2 if split else L.DTYPE_CODE[operand_dtype], L.ptr(wf), L.ptr(c1), L.ptr(c2),
                                           L.stream_ptr(w32.device)), "eoe_vit_fold_layernorm")
    return wf, c1, c2


def gemm_lnfold(A, Wf, c1, c2, stats, quick_gelu=False, shift=None, split=False):
    """rstd*(A @ Wf^T - (mean - shift)*c1) + c2 with (mean, rstd) from the per-row chunk sums `stats` [M, K/128, 2];
    `shift` [M] is what was subtracted from the rows of A (None: nothing)."""
    L.require_cuda(A, Wf, stats)
    M, K = A.shape
    N = Wf.shape[0]
    if split:
        K //= 2
    out = torch.empty(M, (2 if split else 1) * N, dtype=A.dtype, device=A.device)
    # quick_gelu: False / True, or 2 for 1.702 * QuickGELU (EOE_EPI_LNFOLD_QUICKGELU_X1702)
    L.check(L.lib().eoe_gemm_lnfold(L.ptr(A), L.ptr(Wf), L.ptr(c1), L.ptr(c2), L.ptr(stats), L.ptr(shift), L.ptr(out), M, N, K,
                                    L.EOE_F16X2 if split else L.DTYPE_CODE[A.dtype], int(quick_gelu), L.stream_ptr(A.device)),
            "eoe_gemm_lnfold")
    return out


def gemm_residual_stats(A, W, bias, x, stats_in=None, split=False):
    """x += A @ W^T + bias in place; returns (xb, stats, shift): stats [M, N/128, 2] chunk sums of the updated x, shift [M] =
    row means of x BEFORE the update (from `stats_in`, the previous sums; None: zeros), xb = (x - shift) rounded to A.dtype."""
    L.require_cuda(A, W, x)
    M, K = A.shape
    N = W.shape[0]
    if split:
        K //= 2
    xb = torch.empty(M, (2 if split else 1) * N, dtype=A.dtype, device=A.device)
    stats = torch.empty(M, N // 128, 2, dtype=torch.float32, device=A.device)
    shift = torch.empty(M, dtype=torch.float32, device=A.device)
    L.check(L.lib().eoe_gemm_residual_stats(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(stats_in), L.ptr(x), L.ptr(xb), L.ptr(stats),
                                            L.ptr(shift), M, N, K, L.EOE_F16X2 if split else L.DTYPE_CODE[A.dtype],
                                            L.stream_ptr(A.device)),
            "eoe_gemm_residual_stats")
    return xb, stats, shift


def layernorm(x, w, b, out_dtype=torch.bfloat16):
    L.require_cuda(x)
    M, width = x.shape
    split = out_dtype == L.F16X2
    y = torch.empty(M, (2 if split else 1) * width, dtype=torch.float16 if split else out_dtype, device=x.device)
    L.check(L.lib().eoe_layernorm(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), L.EOE_F16X2 if split else L.DTYPE_CODE[out_dtype], M, width,
                                  L.stream_ptr(x.device)), "eoe_layernorm")
    return y


def attention(qkv, B, Lseq, heads, split=False):
    """split=True: qkv [B*L, 6*width] = [q k v | lo halves] and the output [B*L, 2*width] are fp16 (hi | lo) pairs."""
    L.require_cuda(qkv)
    width = qkv.shape[1] // (6 if split else 3)
    out = torch.empty(B * Lseq, (2 if split else 1) * width, dtype=qkv.dtype, device=qkv.device)
    L.check(L.lib().eoe_attention(L.ptr(qkv), L.ptr(out), B, Lseq, heads, L.EOE_F16X2 if split else L.DTYPE_CODE[qkv.dtype],
                                  L.stream_ptr(qkv.device)), "eoe_attention")
    return out
