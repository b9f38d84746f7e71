"""CLIP text tower: `CLIP.encode_text` (clip_official/clip/model.py:339-352) on the same sm_100a kernels as the image
encoder, so that `ADClipTrainer.prepare_metric` (src/eoe/training/clip.py:50-64) needs no PyTorch model at all.

The tower runs once per class on K <= 30 prompts of 77 tokens (2 310 rows), so it is composed from the C ABI's
building blocks, one call per block, instead of a fused plan:
    eoe_text_embed                      token + positional embedding              (model.py:340-342)
    12 x [ eoe_layernorm -> eoe_gemm(bias) -> eoe_attention_causal -> eoe_gemm(+= residual)
           eoe_layernorm -> eoe_gemm(bias, QuickGELU) -> eoe_gemm(+= residual) ]  (model.py:167-188, mask :324-331)
    eoe_text_tail                       ln_final at the <eot> row @ text_projection (model.py:346-350)
Residual stream, LayerNorm statistics and the tail are fp32; GEMM operands fp16 (default: the reference GPU dtype,
clip_official/clip/model.py:371-392), bf16, or "f16x2" (precise mode: fp16 (hi, lo) pairs, 3 MMAs per product, attention in
fp32 -- text features within ~1e-6 of the fp32 reference instead of 7e-4).  Weights load from
the reference's state_dict keys (`token_embedding.weight`, `positional_embedding`, `transformer.resblocks.*`,
`ln_final.*`, `text_projection`).  Tokenisation is host string processing and stays with the caller
(`clip_official/clip/clip.py:164-197` `tokenize`, or any function returning `[K, 77]` int64 ids)."""
from typing import Callable, Dict, Sequence

import torch
from torch import nn

from . import _lib as L
from . import encoder as E


class ClipTextEncoder(nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", operand_dtype=torch.float16, heads: int = None):
        super().__init__()
        self.split = operand_dtype in (L.F16X2, "split")         # precise mode: fp16 (hi, lo) pairs, include/eoe_b200.h EOE_F16X2
        if self.split:
            operand_dtype = torch.float16
        if operand_dtype not in (torch.bfloat16, torch.float16):
            raise L.EoeError('operand_dtype must be torch.bfloat16, torch.float16 or "f16x2"')
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.EoeError("ClipTextEncoder needs a CUDA device (no CPU fallback)")
        sd = state_dict
        self.vocab, self.width = sd["token_embedding.weight"].shape
        self.ctx = sd["positional_embedding"].shape[0]
        self.embed_dim = sd["text_projection"].shape[1]
        self.heads = heads or self.width // 64                       # transformer_heads = transformer_width // 64 (model.py:405)
        self.n_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
        self.operand_dtype = operand_dtype
        self.device_ = dev
        if self.width % 256 != 0 or self.width // self.heads != 64 or self.ctx > 208:
            raise L.EoeError("unsupported text tower: width % 256 == 0, head dim 64 and context <= 208 are required")

        def f32(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def op(t):
            t32 = t.detach().to(device=dev, dtype=torch.float32)
            return E.split_f16(t32) if self.split else t32.to(operand_dtype).contiguous()

        self.tok = f32(sd["token_embedding.weight"])
        self.pos = f32(sd["positional_embedding"])
        self.ln_f = (f32(sd["ln_final.weight"]), f32(sd["ln_final.bias"]))
        self.proj = f32(sd["text_projection"])
        self.blocks = []
        for i in range(self.n_layers):
            p = f"transformer.resblocks.{i}."
            self.blocks.append(dict(
                ln_1=(f32(sd[p + "ln_1.weight"]), f32(sd[p + "ln_1.bias"])),
                ln_2=(f32(sd[p + "ln_2.weight"]), f32(sd[p + "ln_2.bias"])),
                in_w=op(sd[p + "attn.in_proj_weight"]), in_b=f32(sd[p + "attn.in_proj_bias"]),
                out_w=op(sd[p + "attn.out_proj.weight"]), out_b=f32(sd[p + "attn.out_proj.bias"]),
                fc_w=op(sd[p + "mlp.c_fc.weight"]), fc_b=f32(sd[p + "mlp.c_fc.bias"]),
                proj_w=op(sd[p + "mlp.c_proj.weight"]), proj_b=f32(sd[p + "mlp.c_proj.bias"])))

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens [n, ctx] integer ids -> features [n, embed] fp32 (not normalised: prepare_metric does, clip.py:62)."""
        if tokens.dim() != 2 or tokens.shape[1] != self.ctx:
            raise L.EoeError(f"tokens must be [n, {self.ctx}], got {tuple(tokens.shape)}")
        tokens = tokens.detach().to(device=self.device_, dtype=torch.int64).contiguous()
        n = tokens.shape[0]
        if n == 0:
            return torch.empty(0, self.embed_dim, dtype=torch.float32, device=self.device_)
        lo, hi = int(tokens.min()), int(tokens.max())        # nn.Embedding raises on ids outside the table; so do we
        if lo < 0 or hi >= self.vocab:
            raise IndexError(f"token id out of range [0, {self.vocab}): min {lo}, max {hi}")
        lib, st, sp = L.lib(), L.stream_ptr(self.device_), self.split
        dt = L.F16X2 if sp else self.operand_dtype
        x = torch.empty(n * self.ctx, self.width, dtype=torch.float32, device=self.device_)
        L.check(lib.eoe_text_embed(L.ptr(tokens), L.ptr(self.tok), L.ptr(self.pos), L.ptr(x), n, self.ctx, self.width,
                                   self.vocab, st), "eoe_text_embed")
        for b in self.blocks:
            h = E.layernorm(x, *b["ln_1"], out_dtype=dt)
            qkv = E.gemm(h, b["in_w"], b["in_b"], L.EOE_EPI_BIAS, split=sp)
            o = attention_causal(qkv, n, self.ctx, self.heads, split=sp)
            E.gemm(o, b["out_w"], b["out_b"], L.EOE_EPI_BIAS_RESIDUAL_F32, out=x, split=sp)
            h = E.layernorm(x, *b["ln_2"], out_dtype=dt)
            u = E.gemm(h, b["fc_w"], b["fc_b"], L.EOE_EPI_BIAS_QUICKGELU, split=sp)
            E.gemm(u, b["proj_w"], b["proj_b"], L.EOE_EPI_BIAS_RESIDUAL_F32, out=x, split=sp)
        feats = torch.empty(n, self.embed_dim, dtype=torch.float32, device=self.device_)
        L.check(lib.eoe_text_tail(L.ptr(x), L.ptr(tokens), L.ptr(self.ln_f[0]), L.ptr(self.ln_f[1]), L.ptr(self.proj),
                                  L.ptr(feats), n, self.ctx, self.width, self.embed_dim, st), "eoe_text_tail")
        return feats

    encode_text = forward

    def prompt_encoder(self, tokenize: Callable[[Sequence[str]], torch.Tensor]) -> Callable[[Sequence[str]], torch.Tensor]:
        """The `text_encoder` callable ADClipTrainer.prepare_metric expects: prompts -> tokenize -> this tower."""
        return lambda prompts: self(tokenize(list(prompts)))


def attention_causal(qkv, B, Lseq, heads, split=False):
    """split=True: qkv [B*L, 6*width] and the output [B*L, 2*width] are fp16 (hi | lo) pairs (EOE_F16X2)."""
    L.require_cuda(qkv)
    width = qkv.shape[1] // (6 if split else 3)
    out = torch.empty(B * Lseq, (2 if split else 1) * width, dtype=qkv.dtype, device=qkv.device)
    L.check(L.lib().eoe_attention_causal(L.ptr(qkv), L.ptr(out), B, Lseq, heads, L.EOE_F16X2 if split else L.DTYPE_CODE[qkv.dtype],
                                         L.stream_ptr(qkv.device)), "eoe_attention_causal")
    return out
