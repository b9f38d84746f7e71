"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the LIVE reference.

Run in the build container (where /root/reference is mounted):
    python -m oracle.make_golden
The reference has no tests / golden vectors of its own, so these pins
are created from the reference's own code executed here on seeded synthetic inputs:
  heads.npz  HSCTrainer / BCETrainer / ADClipTrainer hooks called unbound
             (src/eoe/training/hsc.py:12-21, bce.py:15-20, clip.py:66-103) + torch autograd grads
  auc_*.npz  sklearn roc_curve/auc/precision_recall_curve/average_precision_score as called at
             src/eoe/training/ad_trainer.py:453-454,517-521
  vit_*.npz  CLIP(...).encode_image (clip_official/clip/model.py:219-236,336-337) with the seeded
             weights of oracle.vit.synth_state_dict loaded into the reference module
  score_parity_b*.npz  encode_image (fp32 and half) -> ADClipTrainer.compute_anomaly_score on 64 images (cfg2 / cfg3)
  text.npz   CLIP(...).encode_text (clip_official/clip/model.py:339-352) with the seeded weights of
             oracle.text.synth_text_state_dict on seeded token rows
Inputs are re-derived from seeds at test time (oracle.golden_inputs) so fixtures stay small;
each file also stores input checksums so generator drift is detected, not silently accepted.
"""
import os
import numpy as np
import torch

from oracle import _ref_import, golden_inputs as gi, text as otext, vit

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Self:
    def __init__(self, ad_mode):
        self.ad_mode = ad_mode


def _grad(fn, z):
    z = z.clone().requires_grad_(True)
    loss = fn(z)
    loss.backward()
    return loss.detach().numpy(), z.grad.numpy()


def make_heads(h):
    out = {}
    z, y = gi.hsc_inputs()
    zt, yt = torch.from_numpy(z), torch.from_numpy(y)
    for nom in (0, 1):
        l, g = _grad(lambda t: h["HSCTrainer"].loss(None, t, yt, None, nominal_label=nom), zt)
        out[f"hsc_loss_nom{nom}"], out[f"hsc_grad_nom{nom}"] = l, g
    out["hsc_score"] = h["HSCTrainer"].compute_anomaly_score(None, zt, None).numpy()
    out["hsc_in_sum"] = np.float64(z.astype(np.float64).sum())

    x, yb = gi.bce_inputs()
    xt, ybt = torch.from_numpy(x), torch.from_numpy(yb)
    l, g = _grad(lambda t: h["BCETrainer"].loss(None, t, ybt, None), xt)
    out["bce_loss"], out["bce_grad"] = l, g
    for nom in (0, 1):
        out[f"bce_score_nom{nom}"] = h["BCETrainer"].compute_anomaly_score(None, xt, None, nominal_label=nom).numpy()
    out["bce_in_sum"] = np.float64(x.astype(np.float64).sum())

    # DSAD / DSVDD / focal (dsad.py:13-22, dsvdd.py:23-27, focal.py:11-39) on the HSC / BCE inputs
    z, y = gi.hsc_inputs()
    zt, yt = torch.from_numpy(z), torch.from_numpy(y)
    for nom in (0, 1):
        l, g = _grad(lambda t: h["DSADTrainer"].loss(None, t, yt, None, nominal_label=nom), zt)
        out[f"dsad_loss_nom{nom}"], out[f"dsad_grad_nom{nom}"] = l, g
    out["dsad_score"] = h["DSADTrainer"].compute_anomaly_score(None, zt, None).numpy()
    zd, cd = gi.dsvdd_inputs()
    zdt, cdt = torch.from_numpy(zd), torch.from_numpy(cd)
    l, g = _grad(lambda t: h["DSVDDTrainer"].loss(None, t, None, cdt), zdt)
    out["dsvdd_loss"], out["dsvdd_grad"] = l, g
    out["dsvdd_score"] = h["DSVDDTrainer"].compute_anomaly_score(None, zdt, cdt).numpy()
    out["dsvdd_in_sum"] = np.float64(zd.astype(np.float64).sum() + cd.astype(np.float64).sum())
    l, g = _grad(lambda t: h["FocalTrainer"].loss(None, t, ybt, None), xt)
    out["focal_loss"], out["focal_grad"] = l, g
    for nom in (0, 1):
        out[f"focal_score_nom{nom}"] = h["FocalTrainer"].compute_anomaly_score(None, xt, None, nominal_label=nom).numpy()

    for K in (2, 10, 30):
        zc, yc, c = gi.clip_inputs(K)
        zt, yt, ct = torch.from_numpy(zc), torch.from_numpy(yc), torch.from_numpy(c)
        out[f"clip_score_K{K}"] = h["ADClipTrainer"].compute_anomaly_score(_Self("leave_one_out"), zt, ct).numpy()
        for mode in ("one_vs_rest", "leave_one_out"):
            for nom in (0, 1):
                l, g = _grad(lambda t: h["ADClipTrainer"].loss(_Self(mode), t, yt, ct, nominal_label=nom), zt)
                out[f"clip_loss_K{K}_{mode}_nom{nom}"] = l
                out[f"clip_grad_K{K}_{mode}_nom{nom}"] = g
        out[f"clip_in_sum_K{K}"] = np.float64(zc.astype(np.float64).sum() + c.astype(np.float64).sum())
    np.savez_compressed(os.path.join(OUT, "heads.npz"), **out)


def make_auc():
    from sklearn.metrics import roc_curve, auc, precision_recall_curve, average_precision_score
    for name in gi.AUC_CASES:
        y, s = gi.auc_inputs(name)
        fpr, tpr, thr = roc_curve(y, s)
        prec, rec, pthr = precision_recall_curve(y, s)
        np.savez_compressed(
            os.path.join(OUT, f"auc_{name}.npz"), auc=np.float64(auc(fpr, tpr)), fpr=fpr, tpr=tpr,
            thresholds=thr.astype(np.float64), ap=np.float64(average_precision_score(y, s)),
            precision=prec, recall=rec, in_sum=np.float64(s.astype(np.float64).sum() + y.sum()))


def make_vit(h):
    for patch in (32, 16):
        sd = vit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
        imgs = gi.vit_images()
        m = h["CLIP"](512, 224, 12, 768, patch, 77, 49408, 512, 8, 12).eval()
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(not k.startswith("visual.") for k in missing)
        with torch.no_grad():
            feats = m.encode_image(imgs).numpy()
        w_sum = float(sum(v.double().sum() for v in sd.values()))
        np.savez_compressed(os.path.join(OUT, f"vit_b{patch}.npz"), features=feats,
                            img_sum=np.float64(imgs.double().sum()), w_sum=np.float64(w_sum))


def live_half_features(h, patch, sd, imgs):
    """The reference's own GPU precision, executed by the live reference on the CPU: `convert_weights` (model.py:371-392)
    + fp16 images (model.py:337).  Rounding points are the reference's; only accumulation order differs from cuBLAS."""
    from eoe.models.clip_official.clip.model import convert_weights
    m = h["CLIP"](512, 224, 12, 768, patch, 77, 49408, 512, 8, 12).eval()
    m.load_state_dict(sd, strict=False)
    convert_weights(m)
    out = []
    with torch.no_grad():
        for s in range(0, imgs.shape[0], 16):
            out.append(m.encode_image(imgs[s:s + 16].half()).float())
    return torch.cat(out).numpy()


def make_score_parity(h):
    """score_parity_b{32,16}.npz: the zero-shot path END TO END through the live reference -- CLIP.encode_image (fp32, and
    in half = the reference's GPU precision) followed by ADClipTrainer.compute_anomaly_score (clip.py:66-79) -- on
    gi.score_parity_inputs.  tests/test_gpu_encoder.py::test_end_to_end_scores compares enc.score() with these."""
    for patch, K in gi.SCORE_PARITY_CFGS:
        sd = vit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
        imgs, text, labels = gi.score_parity_inputs(K)
        m = h["CLIP"](512, 224, 12, 768, patch, 77, 49408, 512, 8, 12).eval()
        m.load_state_dict(sd, strict=False)
        with torch.no_grad():
            f32 = torch.cat([m.encode_image(imgs[s:s + 16]) for s in range(0, imgs.shape[0], 16)])
        f16 = torch.from_numpy(live_half_features(h, patch, sd, imgs))
        tt = torch.from_numpy(text)
        score = lambda f: h["ADClipTrainer"].compute_anomaly_score(_Self("leave_one_out"), f, tt).numpy()
        np.savez_compressed(os.path.join(OUT, f"score_parity_b{patch}.npz"), features=f32.numpy(), scores=score(f32),
                            features_ref_half=f16.numpy(), scores_ref_half=score(f16), text=text, labels=labels,
                            img_sum=np.float64(imgs.double().sum()))


def make_text(h):
    sd = otext.synth_text_state_dict(seed=gi.TEXT_WEIGHT_SEED)
    tokens = gi.text_tokens()
    m = h["CLIP"](512, 224, 2, 768, 32, 77, 49408, 512, 8, 12).eval()       # a 2-layer visual tower: unused here
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("visual.") or k == "logit_scale" for k in missing)
    with torch.no_grad():
        feats = m.encode_text(tokens).numpy()
    w_sum = float(sum(v.double().sum() for v in sd.values()))
    np.savez_compressed(os.path.join(OUT, "text.npz"), features=feats, tok_sum=np.int64(tokens.sum()),
                        w_sum=np.float64(w_sum))


def make_resize():
    """CLIP `_transform` head (clip.py:58-61) run by Pillow + torchvision on seeded images (tests/test_oracle_resize._img)."""
    from PIL import Image
    from torchvision import transforms as T
    import importlib.util
    spec = importlib.util.spec_from_file_location("t", os.path.join(os.path.dirname(OUT), "test_oracle_resize.py"))
    t = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(t)
    tr = T.Compose([T.Resize(224, interpolation=Image.BICUBIC), T.CenterCrop(224)])
    out = {}
    for h, w in [(32, 32), (375, 500), (233, 350)]:
        r = np.asarray(tr(Image.fromarray(t._img(h, w))))
        out[f"sum_{h}x{w}"] = np.int64(r.astype(np.int64).sum())
        out[f"sample_{h}x{w}"] = r[::37, ::41].copy()
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    h = _ref_import.hooks()
    torch.set_num_threads(os.cpu_count())
    make_heads(h)
    make_auc()
    make_vit(h)
    make_score_parity(h)
    make_text(h)
    make_resize()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
