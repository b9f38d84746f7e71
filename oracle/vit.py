"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CLIP ViT-B image encoder forward.

Functional torch-fp32 restatement of the reference's
  VisualTransformer.forward      src/eoe/models/clip_official/clip/model.py:219-236
  ResidualAttentionBlock.forward model.py:167-188 (nn.MultiheadAttention, no mask for vision)
  LayerNorm (fp32, eps 1e-5)     model.py:153-159
  QuickGELU                      model.py:162-164
  CLIP.encode_image              model.py:336-337
operating directly on the reference's state_dict keys (`visual.conv1.weight`,
`visual.transformer.resblocks.{i}.attn.in_proj_weight`, ... model.py:395-402), so the same
weights feed the oracle, the live reference and the CUDA encoder.

Pinned against the live reference `CLIP(...).encode_image` in
tests/test_oracle_vs_reference.py (where /root/reference is mounted) and against
tests/golden/vit_*.npz (features produced by the live reference from seeded weights; generated
by oracle/make_golden.py).

`operand_dtype` (None = exact fp32 oracle) rounds every tensor that the CUDA path stores in
16 bit (GEMM operands: patches, LN outputs, qkv, softmax probabilities, attention output, GELU
output, weights) to that dtype while keeping fp32 accumulation / residual stream / LN / softmax
statistics -- the "precision-matched" oracle used to separate kernel bugs from rounding.

`fold_layernorm=True` (only meaningful with an operand_dtype) moves the rounding points the way the CUDA encoder's
LayerNorm-folded path does (DESIGN.md "LayerNorm fold"): ln_1 / ln_2 are applied as
    rstd * ( round(x) @ round(W*ln_w)^T - mean * c1 ) + c2,   c1 = rowsum(round(W*ln_w)),  c2 = W @ ln_b + bias
with mean / rstd taken from the fp32 residual stream; in exact arithmetic this IS LayerNorm followed by the Linear.
round(x) is really round(x - shift) with shift = the row's mean before its last update (the producers cannot know the new
mean), and the epilogue uses (mean - shift): the 16-bit rounding acts on an (almost) centred row, as LayerNorm's would.
The last block's ln_2 stays unfolded (the CUDA path evaluates it for the class-token rows only).  In the folded blocks
the GELU output is stored as 1.702 * QuickGELU and c_proj's weights are pre-divided by 1.702 (same product).
"""
import math
import torch
import torch.nn.functional as F

N_LAYERS = 12
WIDTH = 768
HEADS = 12
EMBED = 512


def synth_state_dict(patch: int, seed: int = 0, layers: int = N_LAYERS, width: int = WIDTH,
                     embed: int = EMBED, res: int = 224, dtype=torch.float32):
    """Seeded random visual-tower weights with the reference's key names and shapes.
    Scales follow model.py:207-217,295-322 in spirit (std ~ width**-0.5) but are our own
    generator so that the weights are identical on every box (no dependence on nn init order).
    LayerNorm weights/biases are perturbed away from (1,0) so that they are exercised."""
    g = torch.Generator().manual_seed(seed)
    L = (res // patch) ** 2 + 1
    sc = width ** -0.5

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    sd = {
        "visual.conv1.weight": rn(width, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
        "visual.class_embedding": rn(width, std=sc),
        "visual.positional_embedding": rn(L, width, std=sc),
        "visual.ln_pre.weight": 1 + rn(width, std=0.1),
        "visual.ln_pre.bias": rn(width, std=0.1),
        "visual.ln_post.weight": 1 + rn(width, std=0.1),
        "visual.ln_post.bias": rn(width, std=0.1),
        "visual.proj": rn(width, embed, std=sc),
    }
    attn_std = sc
    proj_std = sc * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1 + rn(width, std=0.1)
        sd[p + "ln_1.bias"] = rn(width, std=0.1)
        sd[p + "ln_2.weight"] = 1 + rn(width, std=0.1)
        sd[p + "ln_2.bias"] = rn(width, std=0.1)
        sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
        sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
        sd[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
    return sd


def n_layers_of(sd):
    return len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.resblocks.")})


ROUND_POINTS = ("patch", "w", "xb", "qkv", "p", "h", "u")
_active_points = None          # None = every point; else the subset that is rounded (error attribution, tools/score_parity.py)

# operand_dtype of the CUDA path's precise mode (include/eoe_b200.h EOE_F16X2): every stored 16-bit tensor is an fp16 pair
# hi = rn(x), lo = rn(x - hi) -- the value the kernels multiply with is hi + lo (the lo * lo term they drop is 2^-24
# relative) -- the softmax probabilities included.
F16X2 = "f16x2"


def split_round(t):
    hi = t.to(torch.float16).to(torch.float32)
    return hi + (t - hi).to(torch.float16).to(torch.float32)


def _r(t, dt, point=None):
    if dt is None or (_active_points is not None and point is not None and point not in _active_points):
        return t
    if dt == F16X2:
        return split_round(t)
    return t.to(dt).to(torch.float32)


def _ln_linear(x, ln_w, ln_b, W, bias, dt, fold, shift=None):
    """Linear(LayerNorm(x)) with the rounding points of the selected CUDA path.  `shift` [.., 1]: what the producer
    subtracted from the row before rounding the 16-bit copy (the row's mean before its last update)."""
    width = x.shape[-1]
    if not fold or dt is None:
        h = F.layer_norm(x, (width,), ln_w, ln_b, 1e-5)
        return F.linear(_r(h, dt, "xb"), _r(W, dt, "w"), bias)
    wf = _r(W * ln_w, dt, "w")
    c1 = wf.sum(dim=1)
    c2 = W @ ln_b + bias
    mean = x.mean(dim=-1, keepdim=True)
    var = (x * x).mean(dim=-1, keepdim=True) - mean * mean
    rstd = torch.rsqrt(var.clamp_min(0) + 1e-5)
    if shift is None:
        shift = torch.zeros_like(mean)
    return rstd * (_r(x - shift, dt, "xb") @ wf.t() - (mean - shift) * c1) + c2


@torch.no_grad()
def encode_image(sd, imgs, operand_dtype=None, heads: int = HEADS, return_tokens: bool = False,
                 fold_layernorm: bool = False, points=None):
    """model.py:219-236. imgs [B,3,R,R] float32 (already CLIP-normalised) -> features [B, embed].
    points: subset of ROUND_POINTS that is rounded to operand_dtype (None = all of them)."""
    global _active_points
    _active_points = None if points is None else set(points)
    try:
        return _encode_image(sd, imgs, operand_dtype, heads, return_tokens, fold_layernorm)
    finally:
        _active_points = None


def _encode_image(sd, imgs, operand_dtype, heads, return_tokens, fold_layernorm):
    dt = operand_dtype
    f32 = torch.float32
    w = {k: v.to(f32) for k, v in sd.items()}
    conv_w = w["visual.conv1.weight"]
    width, _, P, _ = conv_w.shape
    B = imgs.shape[0]
    layers = n_layers_of(sd)
    # model.py:220-222  conv1 (stride = kernel = P, no bias) -> [B, g*g, width]
    x = F.conv2d(_r(imgs.to(f32), dt, "patch"), _r(conv_w, dt, "w"), stride=P)
    x = x.reshape(B, width, -1).permute(0, 2, 1)
    # model.py:223-225  class token, positional embedding, ln_pre
    cls = w["visual.class_embedding"].expand(B, 1, width)
    x = torch.cat([cls, x], dim=1) + w["visual.positional_embedding"]
    x = F.layer_norm(x, (width,), w["visual.ln_pre.weight"], w["visual.ln_pre.bias"], 1e-5)
    L = x.shape[1]
    dh = width // heads
    shift = x.mean(dim=-1, keepdim=True)          # ln_pre centres its 16-bit copy on the row's own mean
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        # model.py:186  x = x + attn(ln_1(x))
        qkv = _ln_linear(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"], w[p + "attn.in_proj_weight"],
                         w[p + "attn.in_proj_bias"], dt, fold_layernorm, shift)
        qkv = _r(qkv, dt, "qkv")
        q, k, v = qkv.split(width, dim=-1)
        q = q.reshape(B, L, heads, dh).transpose(1, 2)
        k = k.reshape(B, L, heads, dh).transpose(1, 2)
        v = v.reshape(B, L, heads, dh).transpose(1, 2)
        s = None if dt is None else (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
        if dt is None:
            # fused attention, as nn.MultiheadAttention runs it in the live reference (same math; keeps the CPU baseline
            # this oracle doubles as -- bench.py cpu_baseline -- as fast as the reference itself, tools/anchor_cpu_arm.py)
            o = F.scaled_dot_product_attention(q, k, v)
        else:
            # CUDA path: un-normalised exp in 16 bit for the PV product, fp32 row sum, divide after
            e = torch.exp(s - s.amax(dim=-1, keepdim=True))
            o = (_r(e, dt, "p") @ v) / e.sum(dim=-1, keepdim=True)
        o = _r(o.transpose(1, 2).reshape(B, L, width), dt, "h")
        shift = x.mean(dim=-1, keepdim=True)      # the residual GEMM centres xb on the row's mean BEFORE its update
        x = x + F.linear(o, _r(w[p + "attn.out_proj.weight"], dt, "w"), w[p + "attn.out_proj.bias"])
        # model.py:187  x = x + mlp(ln_2(x)),  mlp = c_proj(QuickGELU(c_fc(.)))  (model.py:173-177)
        fold_mlp = fold_layernorm and dt is not None and i < layers - 1
        u = _ln_linear(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"], w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"], dt,
                       fold_mlp, shift)
        shift = x.mean(dim=-1, keepdim=True)
        if fold_mlp:
            # the folded path stores 1.702 * QuickGELU and multiplies by c_proj weights pre-divided by 1.702
            u = _r(1.702 * u * torch.sigmoid(1.702 * u), dt, "u")
            x = x + F.linear(u, _r(w[p + "mlp.c_proj.weight"] / 1.702, dt, "w"), w[p + "mlp.c_proj.bias"])
        else:
            u = _r(u * torch.sigmoid(1.702 * u), dt, "u")
            x = x + F.linear(u, _r(w[p + "mlp.c_proj.weight"], dt, "w"), w[p + "mlp.c_proj.bias"])
    if return_tokens:
        return x
    # model.py:231-234  ln_post on the class token, then @ proj
    c = F.layer_norm(x[:, 0, :], (width,), w["visual.ln_post.weight"], w["visual.ln_post.bias"], 1e-5)
    return c @ w["visual.proj"]


@torch.no_grad()
def encode_image_ref_fp16(sd, imgs, heads: int = HEADS):
    """Emulation of the reference's OWN GPU precision (model.py:371-392 `convert_weights`: Conv / Linear /
    MultiheadAttention weights + biases and `proj` in fp16; clip_official/clip/clip.py:115-116 keeps fp32 only on CPU):
    the image is cast to fp16 (model.py:337), class / positional embeddings are cast to the activation dtype
    (model.py:222-223), so the residual stream and every activation are fp16 tensors; LayerNorm computes in fp32 from
    the fp16 input and casts back (model.py:156-159).  Arithmetic INSIDE an op is fp32 here (cuBLAS / SDPA accumulate in
    fp32), each op's output is rounded to fp16 -- the favourable reading of that path.  Pinned against the live reference
    run in half on the CPU (tests/test_oracle_vs_reference.py::test_ref_fp16_emulation).  Used to bound our 16-bit
    paths: "no further from the fp32 answer than the reference's own GPU path"."""
    h16, f32 = torch.float16, torch.float32

    def r(t):
        return t.to(h16).to(f32)

    w = {k: v.to(f32) for k, v in sd.items()}
    for k in w:
        if k.endswith(("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "c_fc.weight", "c_fc.bias",
                       "c_proj.weight", "c_proj.bias", "conv1.weight")) or k == "visual.proj":
            w[k] = r(w[k])

    def ln(x, name):
        return r(F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], 1e-5))

    conv_w = w["visual.conv1.weight"]
    width, _, P, _ = conv_w.shape
    B = imgs.shape[0]
    layers = n_layers_of(sd)
    x = r(F.conv2d(r(imgs.to(f32)), conv_w, stride=P))
    x = x.reshape(B, width, -1).permute(0, 2, 1)
    x = torch.cat([r(w["visual.class_embedding"]).expand(B, 1, width), x], dim=1)
    x = r(x + r(w["visual.positional_embedding"]))
    x = ln(x, "visual.ln_pre")
    L = x.shape[1]
    dh = width // heads
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        qkv = r(F.linear(ln(x, p + "ln_1"), w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"]))
        q, k, v = qkv.split(width, dim=-1)
        q = q.reshape(B, L, heads, dh).transpose(1, 2)
        k = k.reshape(B, L, heads, dh).transpose(1, 2)
        v = v.reshape(B, L, heads, dh).transpose(1, 2)
        s = r((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh)))
        o = r(r(s.softmax(dim=-1)) @ v)
        o = o.transpose(1, 2).reshape(B, L, width)
        x = r(x + r(F.linear(o, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"])))
        u = r(F.linear(ln(x, p + "ln_2"), w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]))
        u = r(u * r(torch.sigmoid(r(1.702 * u))))
        x = r(x + r(F.linear(u, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])))
    return r(ln(x[:, 0, :], "visual.ln_post") @ w["visual.proj"])


def flops_per_image(patch: int, res: int = 224, layers: int = N_LAYERS, width: int = WIDTH,
                    embed: int = EMBED) -> float:
    """2*MAC over GEMMs + QK^T + PV: 8.818e9 (B/32), 35.127e9 (B/16)."""
    g2 = (res // patch) ** 2
    L = g2 + 1
    pe = 2 * g2 * 3 * patch * patch * width
    per_layer = 2 * L * width * (3 * width + width + 4 * width + 4 * width) + 2 * 2 * L * L * width
    return float(pe + layers * per_layer + 2 * width * embed)
