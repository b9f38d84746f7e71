"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CLIP text tower, `CLIP.encode_text`
(reference src/eoe/models/clip_official/clip/model.py:339-352; causal mask :324-331; blocks :167-188), which
`ADClipTrainer.prepare_metric` (src/eoe/training/clip.py:50-64) runs once per class on the tokenised prompts.
Never imported by the product path (eoe_b200/); pinned against the live reference in tests/test_oracle_vs_reference.py
and against tests/golden/text.npz (features of the live reference, oracle/make_golden.py).

State-dict keys are the reference's: `token_embedding.weight [V, W]`, `positional_embedding [ctx, W]`,
`transformer.resblocks.{i}.{ln_1,ln_2}.{weight,bias}`, `.attn.in_proj_{weight,bias}`, `.attn.out_proj.{weight,bias}`,
`.mlp.c_fc.{weight,bias}`, `.mlp.c_proj.{weight,bias}`, `ln_final.{weight,bias}`, `text_projection [W, E]`.

`operand_dtype` rounds exactly the tensors the CUDA path (eoe_b200.text_encoder.ClipTextEncoder) stores in 16 bit: LayerNorm
outputs, weights, qkv, probabilities, attention output, GELU output; residual stream, LayerNorm / softmax statistics,
token + positional embedding, ln_final and the projection stay fp32.
"""
import torch
import torch.nn.functional as F

TEXT_WIDTH = 512
TEXT_HEADS = 8
TEXT_LAYERS = 12
CTX = 77
VOCAB = 49408
EMBED = 512


def synth_text_state_dict(seed: int = 0, layers: int = TEXT_LAYERS, width: int = TEXT_WIDTH, embed: int = EMBED,
                          ctx: int = CTX, vocab: int = VOCAB):
    """Seeded random text-tower weights with the reference's key names and shapes; scales as model.py:295-322
    (token embedding std 0.02, positional 0.01, attn width**-0.5, proj (2*layers)**-0.5 smaller, fc (2*width)**-0.5)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sc = width ** -0.5
    sd = {
        "token_embedding.weight": rn(vocab, width, std=0.02),
        "positional_embedding": rn(ctx, width, std=0.01),
        "ln_final.weight": 1 + rn(width, std=0.1),
        "ln_final.bias": rn(width, std=0.1),
        "text_projection": rn(width, embed, std=sc),
    }
    proj_std = sc * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1 + rn(width, std=0.1)
        sd[p + "ln_1.bias"] = rn(width, std=0.1)
        sd[p + "ln_2.weight"] = 1 + rn(width, std=0.1)
        sd[p + "ln_2.bias"] = rn(width, std=0.1)
        sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=sc)
        sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
        sd[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
    return sd


def synth_tokens(n: int, seed: int = 0, ctx: int = CTX, vocab: int = VOCAB, min_len: int = 3):
    """Token rows shaped like `tokenize` output (clip_official/clip/clip.py:164-197): <sot> = vocab-2, body tokens,
    <eot> = vocab-1 (the largest id, which encode_text's argmax looks for), zero padding.  Row 0 is as short as a prompt
    can be, row n-1 fills the whole context."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n, ctx, dtype=torch.int64)
    for i in range(n):
        ln = int(torch.randint(min_len, ctx - 2, (1,), generator=g))
        if i == 0:
            ln = 0
        if i == n - 1:
            ln = ctx - 2
        t[i, 0] = vocab - 2
        if ln:
            t[i, 1:1 + ln] = torch.randint(1, vocab - 2, (ln,), generator=g)
        t[i, 1 + ln] = vocab - 1
    return t


def n_layers_of(sd):
    return len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})


def _r(t, dt, is_prob=False):
    if dt is None:
        return t
    if dt == "f16x2":          # precise mode (oracle.vit.F16X2): fp16 (hi, lo) pairs; the text tower's attention keeps P in fp32
        if is_prob:
            return t
        hi = t.to(torch.float16).to(torch.float32)
        return hi + (t - hi).to(torch.float16).to(torch.float32)
    return t.to(dt).to(torch.float32)


@torch.no_grad()
def encode_text(sd, tokens, operand_dtype=None, heads: int = TEXT_HEADS):
    """tokens [n, ctx] int64 -> [n, embed] fp32 (not normalised; prepare_metric does that, clip.py:62)."""
    dt = operand_dtype
    w = {k: v.to(torch.float32) for k, v in sd.items() if not k.startswith("visual.")}
    layers = n_layers_of(w)
    width = w["ln_final.weight"].numel()
    n, ctx = tokens.shape
    dh = width // heads
    # model.py:340-342  token embedding + positional embedding
    x = w["token_embedding.weight"][tokens] + w["positional_embedding"][:ctx]
    # model.py:324-331  additive causal mask: -inf strictly above the diagonal
    mask = torch.full((ctx, ctx), float("-inf")).triu_(1)
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        h = _r(F.layer_norm(x, (width,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], 1e-5), dt)
        qkv = _r(F.linear(h, _r(w[p + "attn.in_proj_weight"], dt), w[p + "attn.in_proj_bias"]), dt)
        q, k, v = qkv.split(width, dim=-1)
        q = q.reshape(n, ctx, heads, dh).transpose(1, 2)
        k = k.reshape(n, ctx, heads, dh).transpose(1, 2)
        v = v.reshape(n, ctx, heads, dh).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / (dh ** 0.5) + mask
        if dt is None:
            o = torch.softmax(s, dim=-1) @ v
        else:
            e = torch.exp(s - s.amax(dim=-1, keepdim=True))
            o = (_r(e, dt, is_prob=True) @ v) / e.sum(dim=-1, keepdim=True)
        o = _r(o.transpose(1, 2).reshape(n, ctx, width), dt)
        x = x + F.linear(o, _r(w[p + "attn.out_proj.weight"], dt), w[p + "attn.out_proj.bias"])
        h = _r(F.layer_norm(x, (width,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], 1e-5), dt)
        u = F.linear(h, _r(w[p + "mlp.c_fc.weight"], dt), w[p + "mlp.c_fc.bias"])
        u = _r(u * torch.sigmoid(1.702 * u), dt)
        x = x + F.linear(u, _r(w[p + "mlp.c_proj.weight"], dt), w[p + "mlp.c_proj.bias"])
    # model.py:346-350  ln_final, row of the <eot> token (largest id; first occurrence), @ text_projection
    eot = tokens.argmax(dim=-1)
    c = F.layer_norm(x[torch.arange(n), eot], (width,), w["ln_final.weight"], w["ln_final.bias"], 1e-5)
    return c @ w["text_projection"]
