"""Random-init CLIP visual-tower weights with the reference's state_dict keys / shapes
(clip_official/clip/model.py:207-217,395-402) for benchmarks: no network, so no pretrained checkpoint."""
import torch


def random_vit_state_dict(patch: int, seed: int = 0, layers: int = 12, width: int = 768, embed: int = 512,
                          res: int = 224, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = (res // patch) ** 2 + 1
    sc = width ** -0.5

    def rn(*shape, std=1.0, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).to(device)

    sd = {
        "visual.conv1.weight": rn(width, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
        "visual.class_embedding": rn(width, std=sc),
        "visual.positional_embedding": rn(L, width, std=sc),
        "visual.ln_pre.weight": rn(width, std=0.1, mean=1.0), "visual.ln_pre.bias": rn(width, std=0.1),
        "visual.ln_post.weight": rn(width, std=0.1, mean=1.0), "visual.ln_post.bias": rn(width, std=0.1),
        "visual.proj": rn(width, embed, std=sc),
    }
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = rn(width, std=0.1, mean=1.0), rn(width, std=0.1)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = rn(width, std=0.1, mean=1.0), rn(width, std=0.1)
        sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = rn(3 * width, width, std=sc), rn(3 * width, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(width, width, std=sc * (2 * layers) ** -0.5)
        sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
        sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = rn(4 * width, width, std=(2 * width) ** -0.5), rn(4 * width, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=sc * (2 * layers) ** -0.5)
        sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
    return sd
