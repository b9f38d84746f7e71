"""GPU: the trainer plug-ins behave like the reference's hooks inside the re-hosted train/eval loops.
The comparison arm is a plain torch statement of the reference objective (hsc.py:17-21 / bce.py:19-20) driving the
same model from the same initial weights with the same optimiser; losses must agree step for step."""
import copy

import numpy as np
import pytest
import torch

from oracle import auc as oauc
from oracle import heads as oh

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _loader(n_batches, d_in, seed, clf=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        imgs = torch.randn(256, d_in, generator=g)
        imgs[128:] += 0.7                                      # "OE" half is shifted
        lbls = torch.cat([torch.zeros(128, dtype=torch.long), torch.ones(128, dtype=torch.long)])   # bases.py:591-597
        out.append((imgs, lbls, torch.arange(i * 256, (i + 1) * 256)))
    return out


def _mlp(d_in, d_out):
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(d_in, 128), torch.nn.LeakyReLU(), torch.nn.Linear(128, d_out))


def _ref_hsc(z, y, nominal=0):
    d = torch.sqrt(torch.norm(z, p=2, dim=1) ** 2 + 1) - 1
    s = 1 - torch.exp(-d)
    return torch.where(y == nominal, d, -torch.log(s + 1e-9)).mean()


def _ref_focal(x, y, gamma=2.0, eps=1e-7):                 # focal.py:19-24
    b = torch.nn.functional.binary_cross_entropy_with_logits(x, y, reduction="none")
    pt = torch.exp(-b).clamp(eps, 1.0 - eps)
    return ((1 - pt).pow(gamma) * b).mean()


def _ref_dsvdd_center(model, loader, eps=1e-1):            # dsvdd.py:11-21
    center = []
    for imgs, lbls, _ in loader:
        with torch.no_grad():
            center.append(model(imgs.to(DEV)[lbls.to(DEV) == 0]).cpu().mean(0).unsqueeze(0))
    center = torch.cat(center).mean(0).unsqueeze(0).to(DEV)
    center[(abs(center) < eps) & (center < 0)] = -eps
    center[(abs(center) < eps) & (center > 0)] = eps
    return center


@pytest.mark.parametrize("objective", ["hsc", "bce", "dsad", "dsvdd", "focal"])
def test_training_loop_matches_reference_objective(objective):
    from eoe_b200.training import TRAINER
    d_out = 1 if objective in ("bce", "focal") else 64
    loader = _loader(4, 32, seed=3)
    model = _mlp(32, d_out)
    ref_model = copy.deepcopy(model).to(DEV)
    tr = TRAINER[objective](model, epochs=2, lr=1e-3, wdk=0.0, milestones=[1], batch_size=128, device=DEV)
    model, roc, losses = tr.train_cls(model, loader, nominal_label=0, clsstr="airplane")
    # reference arm
    opt = torch.optim.Adam(ref_model.parameters(), lr=1e-3, weight_decay=0.0)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [1], 0.1)
    ref_losses, last_scores, last_labels = [], [], []
    ref_center = _ref_dsvdd_center(ref_model, loader) if objective == "dsvdd" else None
    if ref_center is not None:
        torch.testing.assert_close(tr.center, ref_center, rtol=1e-5, atol=1e-6)
    for ep in range(2):
        acc, last_scores, last_labels = [], [], []
        for imgs, lbls, _ in loader:
            imgs, lbls = imgs.to(DEV), lbls.to(DEV)
            opt.zero_grad()
            f = ref_model(imgs)
            if objective == "hsc":
                loss = _ref_hsc(f, lbls)
                sc = 1 - torch.exp(-(torch.sqrt(torch.norm(f, p=2, dim=1) ** 2 + 1) - 1))
            elif objective == "dsad":                                  # dsad.py:13-22
                d2 = torch.norm(f, p=2, dim=1) ** 2
                loss = torch.where(lbls == 0, d2, (d2 + 1e-9) ** (-1)).mean()
                sc = 1 - torch.exp(-(torch.sqrt(d2 + 1) - 1))
            elif objective == "dsvdd":                                 # dsvdd.py:23-27
                sc = (f - ref_center).pow(2).sum(-1)
                loss = sc.mean()
            elif objective == "focal":
                loss = _ref_focal(f.squeeze(), lbls.float())
                sc = torch.sigmoid(f).squeeze()
            else:
                loss = torch.nn.functional.binary_cross_entropy_with_logits(f.squeeze(), lbls.float())
                sc = torch.sigmoid(f).squeeze()
            loss.backward()
            opt.step()
            acc.append(loss.item())
            last_scores.append(sc.detach().cpu())
            last_labels.append(lbls.cpu())
        ref_losses.append(float(np.mean(acc)))
        sched.step()
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-3)          # north_star tolerance for losses
    for p, q in zip(model.parameters(), ref_model.parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-3, atol=2e-5)
    want_auc = oauc.roc_auc(torch.cat(last_labels).numpy(), torch.cat(last_scores).numpy())
    assert abs(roc.auc - want_auc) < 1e-3                               # scores differ in the last ulp -> ties may move


def test_eval_loop_returns_reference_containers():
    from eoe_b200.training import HSCTrainer
    from sklearn.metrics import average_precision_score, roc_auc_score
    model = _mlp(32, 64).to(DEV)
    tr = HSCTrainer(model, device=DEV)
    tr.center = None
    loader = _loader(3, 32, seed=5)
    loader[1][1][5] = -1                                                # unlabeled sample is ignored (ad_trainer.py:517)
    roc, prc = tr.eval_cls(model, loader, nominal_label=0)
    labels, scores = tr.last_eval
    keep = labels.cpu().numpy() >= 0
    y, s = labels.cpu().numpy()[keep], scores.cpu().numpy()[keep]
    assert roc.auc == roc_auc_score(y, s) and roc.auc == oauc.roc_auc(y, s)     # AUC bit-exact given identical scores
    assert prc.avg_prec == average_precision_score(y, s)
    assert roc.fpr[0] == 0 and roc.tpr[-1] == 1 and roc.ths[0] == np.inf
    np.testing.assert_allclose(s, oh.hsc_score(model(torch.cat([b[0] for b in loader]).to(DEV)).detach().cpu().numpy()[keep]),
                               rtol=1e-3, atol=1e-7)


def test_nan_scores_raise_like_the_reference():
    from eoe_b200.training import HSCTrainer, NanGradientsError
    model = _mlp(32, 64)
    with torch.no_grad():
        model[2].weight[0, 0] = float("nan")
    tr = HSCTrainer(model, epochs=1, device=DEV)
    with pytest.raises(NanGradientsError):
        tr.train_cls(model, _loader(1, 32, seed=1))


def test_clip_trainer_zero_shot_eval():
    """epochs == 0 => zero-shot (ad_trainer.py:406 never runs): prepare_metric, then eval through the B200 encoder."""
    from eoe_b200.encoder import ClipImageEncoder
    from eoe_b200.synth import random_vit_state_dict
    from eoe_b200.training import ADClipTrainer
    enc = ClipImageEncoder(random_vit_state_dict(32, seed=2, layers=2), device=DEV, max_batch=16)
    g = torch.Generator().manual_seed(7)
    text = torch.randn(2, 512, generator=g)
    tr = ADClipTrainer(enc, device=DEV, text_features={"airplane": text}, epochs=0)
    _, roc0, losses = tr.train_cls(enc, [], clsstr="airplane")
    assert roc0 is None and losses == []
    assert torch.allclose(tr.center.norm(dim=-1).cpu(), torch.ones(2))
    batches = [(torch.randn(16, 3, 224, 224, generator=g), (torch.rand(16, generator=g) < 0.5).long(), torch.arange(16))
               for _ in range(2)]
    roc, prc = tr.eval_cls(enc, batches, nominal_label=0)
    labels, scores = tr.last_eval
    feats = enc(torch.cat([b[0] for b in batches]).to(DEV)).cpu().numpy()
    np.testing.assert_allclose(scores.cpu().numpy(), oh.clip_score(feats, tr.center.cpu().numpy()), rtol=1e-3, atol=1e-30)
    assert roc.auc == oauc.roc_auc(labels.cpu().numpy(), scores.cpu().numpy())


def test_clip_trainer_zero_shot_eval_precise_mode_matches_the_fp32_reference_path():
    """cfg2 through the plug-in API, end to end, in the precise mode: ClipModel (both towers, operand_dtype "f16x2") behind
    ADClipTrainer -- prompts' token ids -> prepare_metric -> eval_cls -> scores + AUC -- against the fp32 ORACLE of the whole
    reference path (oracle.text -> normalise -> oracle.vit -> oracle.heads.clip_score -> oracle.auc): every score within
    north_star's 1e-3 relative (measured ~2e-5) and the AUC identical."""
    from eoe_b200.clip_model import ClipModel
    from eoe_b200.training import ADClipTrainer
    from oracle import golden_inputs as gi, text as otext, vit as ovit
    sd = {**ovit.synth_state_dict(32, seed=gi.VIT_WEIGHT_SEED), **otext.synth_text_state_dict(seed=gi.TEXT_WEIGHT_SEED)}
    m = ClipModel(sd, device=DEV, operand_dtype="f16x2", max_batch=16)
    tok = gi.text_tokens()                                              # 10 prompts (9 classes + the anomaly prompt)
    seen = []

    def text_encoder(prompts):                                          # the caller's tokenizer: here the seeded ids, one row per prompt
        seen.append(list(prompts))
        assert len(prompts) == tok.shape[0]
        return m.encode_text(tok.to(DEV))

    tr = ADClipTrainer(m, device=DEV, epochs=0, ad_mode="leave_one_out", text_encoder=text_encoder,
                       class_names=[f"c{i}" for i in range(10)])       # leave_one_out: 9 nominal prompts + the anomaly prompt
    tr.train_cls(m, [], clsstr="c3")
    assert seen and seen[0][-1] == "a photo of something" and "a photo of a c3" not in seen[0]
    g = torch.Generator().manual_seed(17)
    batches = [(torch.randn(16, 3, 224, 224, generator=g), (torch.rand(16, generator=g) < 0.5).long(), torch.arange(16))
               for _ in range(2)]
    roc, _ = tr.eval_cls(m, batches, nominal_label=0)
    labels, scores = tr.last_eval
    t32 = otext.encode_text(sd, tok)
    c32 = (t32 / t32.norm(dim=-1, keepdim=True)).numpy()
    torch.testing.assert_close(tr.center.cpu(), torch.from_numpy(c32), rtol=0, atol=2e-6)
    want = oh.clip_score(ovit.encode_image(sd, torch.cat([b[0] for b in batches])).numpy(), c32)
    s = scores.cpu().numpy().astype(np.float64)
    rel = np.abs(s - want) / np.abs(want)
    assert rel.max() <= 1e-3 and rel.max() <= 2e-4, (float(np.median(rel)), float(rel.max()))
    assert roc.auc == oauc.roc_auc(labels.cpu().numpy(), want.astype(np.float32))


def test_cfg5_hsc_on_clip_vitb16_features_with_sgd():
    """BASELINE config 5, the part on the hot path: [n, 512] features of the B200 ViT-B/16 tower (frozen: the encoder is
    forward-only) -> HSC loss + backward + scores, optimiser = SGD(nesterov, momentum 0.9) as the reference picks for CLIP
    models (ad_trainer.py:379-381).  A trainable head on top of the features is driven by both arms."""
    from eoe_b200.encoder import ClipImageEncoder
    from eoe_b200.synth import random_vit_state_dict
    from eoe_b200.training import HSCTrainer
    enc = ClipImageEncoder(random_vit_state_dict(16, seed=4, layers=2), device=DEV, max_batch=32)
    g = torch.Generator().manual_seed(11)
    imgs = torch.randn(64, 3, 224, 224, generator=g)
    with torch.no_grad():
        feats = enc(imgs.to(DEV)).cpu()                               # [64, 512] image features
    assert feats.shape == (64, 512) and torch.isfinite(feats).all()
    lbls = torch.cat([torch.zeros(32, dtype=torch.long), torch.ones(32, dtype=torch.long)])
    loader = [(feats[i:i + 32] * 0.2, lbls[torch.arange(i, i + 32) % 64].clone(), torch.arange(i, i + 32)) for i in (0, 32)]
    loader[0][1][16:] = 1                                             # both classes in every batch
    loader[1][1][:16] = 0
    torch.manual_seed(0)
    head = torch.nn.Linear(512, 512)
    ref_head = copy.deepcopy(head).to(DEV)
    tr = HSCTrainer(head, epochs=3, lr=1e-2, wdk=1e-4, milestones=[2], batch_size=32, device=DEV, sgd=True)
    head, roc, losses = tr.train_cls(head, loader, nominal_label=0, clsstr="acorn")
    opt = torch.optim.SGD(ref_head.parameters(), lr=1e-2, weight_decay=1e-4, momentum=0.9, nesterov=True)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [2], 0.1)
    ref_losses = []
    for ep in range(3):
        acc = []
        for x, y, _ in loader:
            x, y = x.to(DEV), y.to(DEV)
            opt.zero_grad()
            loss = _ref_hsc(ref_head(x), y)
            loss.backward()
            opt.step()
            acc.append(loss.item())
        ref_losses.append(float(np.mean(acc)))
        sched.step()
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-3)
    for p, q in zip(head.parameters(), ref_head.parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-3, atol=2e-5)
    assert roc is not None and 0.0 <= roc.auc <= 1.0


def _cnn32_like():
    """the shape of the reference's CIFAR network (models/cnn.py:44-86: 3 x [conv5x5 - BatchNorm - LeakyReLU - pool],
    two linear layers, rep_dim 256, BatchNorm everywhere) restated small; the reference model itself stays out of the product"""
    torch.manual_seed(0)
    nn = torch.nn
    return nn.Sequential(
        nn.Conv2d(3, 32, 5, padding=2, bias=False), nn.BatchNorm2d(32, eps=1e-4, affine=False), nn.LeakyReLU(), nn.MaxPool2d(2),
        nn.Conv2d(32, 64, 5, padding=2, bias=False), nn.BatchNorm2d(64, eps=1e-4, affine=False), nn.LeakyReLU(), nn.MaxPool2d(2),
        nn.Conv2d(64, 128, 5, padding=2, bias=False), nn.BatchNorm2d(128, eps=1e-4, affine=False), nn.LeakyReLU(), nn.MaxPool2d(2),
        nn.Flatten(), nn.Linear(128 * 4 * 4, 512, bias=False), nn.BatchNorm1d(512, eps=1e-4, affine=False), nn.LeakyReLU(),
        nn.Linear(512, 256, bias=False))


@pytest.mark.parametrize("objective,sgd", [("hsc", False), ("bce", False), ("hsc", True)])
def test_graph_captured_step_equals_eager(objective, sgd):
    """graph_step=True replays one captured CUDA graph per batch (model forward, fused loss + score kernel, backward,
    optimiser, BatchNorm statistics, lr schedule through a device tensor): same losses, scores and weights as the eager
    loop, cfg1's shapes (256 x 3 x 32 x 32 per batch, CNN with BatchNorm), including a ragged last batch run eagerly."""
    from eoe_b200.training import TRAINER
    g = torch.Generator().manual_seed(5)
    loader = []
    for i in range(5):
        nb = 256 if i < 4 else 96
        imgs = torch.randn(nb, 3, 32, 32, generator=g)
        lbls = (torch.arange(nb) >= nb // 2).long()
        imgs[lbls == 1] += 0.5
        loader.append((imgs, lbls, torch.arange(nb)))
    res = {}
    for graph in (False, True):
        model = _cnn32_like()
        if objective == "bce":
            model.append(torch.nn.Linear(256, 1))
        tr = TRAINER[objective](model, epochs=3, lr=1e-3 if not sgd else 1e-2, milestones=[2], device=DEV, graph_step=graph, sgd=sgd)
        model, roc, losses = tr.train_cls(model, loader, nominal_label=0)
        res[graph] = (losses, roc.auc, [p.detach().clone() for p in model.parameters()],
                      [b.detach().clone().float() for b in model.buffers()])
    # The first epoch agrees to rounding (same arithmetic, replayed); afterwards the two runs drift apart like any two
    # runs of this BatchNorm CNN whose convolutions were scheduled differently (cuDNN picks algorithms per call; the
    # differences in the last bit are amplified by training), and Adam moves weights whose gradient is rounding noise by
    # +-lr per step whatever its size -- so later epochs are held to a few percent, final AUC to 2e-2.
    np.testing.assert_allclose(res[True][0][0], res[False][0][0], rtol=2e-3)
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=5e-2)
    assert abs(res[True][1] - res[False][1]) < 2e-2
    if sgd:
        for a, b in zip(res[True][2], res[False][2]):
            torch.testing.assert_close(a, b, rtol=0.1, atol=2e-2)
