"""CPU, build container only: oracle restatements vs the LIVE reference imported from
/root/reference (skipped on the GPU box where the reference is not mounted)."""
import numpy as np
import pytest
import torch

from oracle import _ref_import
from oracle import golden_inputs as gi
from oracle import heads as oh
from oracle import text as otext
from oracle import vit as ovit

pytestmark = pytest.mark.reference


class _Self:
    def __init__(self, ad_mode):
        self.ad_mode = ad_mode


@pytest.fixture(scope="module")
def ref():
    return _ref_import.hooks()


@pytest.mark.parametrize("n,d", [(1, 8), (5, 33), (256, 256), (64, 512)])
def test_hsc_live(ref, n, d):
    rng = np.random.default_rng(n + d)
    z = (0.05 * rng.standard_normal((n, d))).astype(np.float32)
    y = rng.integers(0, 2, n)
    zt = torch.from_numpy(z).requires_grad_(True)
    for nom in (0, 1):
        loss = ref["HSCTrainer"].loss(None, zt, torch.from_numpy(y), None, nominal_label=nom)
        zt.grad = None
        loss.backward()
        np.testing.assert_allclose(oh.hsc_loss(z, y, nom), loss.item(), rtol=1e-5)
        np.testing.assert_allclose(oh.hsc_grad(z, y, nom), zt.grad.numpy(), rtol=3e-5, atol=1e-9)
    sc = ref["HSCTrainer"].compute_anomaly_score(None, zt.detach(), None).numpy()
    np.testing.assert_allclose(oh.hsc_score(z), sc, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("n", [1, 2, 257])
def test_bce_live(ref, n):
    if n == 1:
        pytest.skip("features.squeeze() of [1,1] is 0-dim; the reference errors on labels [1]")
    rng = np.random.default_rng(n)
    x = (4 * rng.standard_normal((n, 1))).astype(np.float32)
    y = rng.integers(0, 2, n)
    xt = torch.from_numpy(x).requires_grad_(True)
    loss = ref["BCETrainer"].loss(None, xt, torch.from_numpy(y), None)
    loss.backward()
    np.testing.assert_allclose(oh.bce_loss(x, y), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(oh.bce_grad(x, y), xt.grad.numpy().reshape(-1), rtol=1e-5, atol=1e-10)
    for nom in (0, 1):
        sc = ref["BCETrainer"].compute_anomaly_score(None, xt.detach(), None, nominal_label=nom).numpy()
        np.testing.assert_allclose(oh.bce_score(x, nom), sc, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("n,d", [(2, 8), (5, 33), (256, 256)])
def test_dsad_dsvdd_live(ref, n, d):
    rng = np.random.default_rng(n * d)
    z = (0.3 * rng.standard_normal((n, d))).astype(np.float32)
    y = rng.integers(0, 2, n)
    c = (0.2 * rng.standard_normal((1, d))).astype(np.float32)
    zt = torch.from_numpy(z).requires_grad_(True)
    for nom in (0, 1):
        loss = ref["DSADTrainer"].loss(None, zt, torch.from_numpy(y), None, nominal_label=nom)
        zt.grad = None
        loss.backward()
        np.testing.assert_allclose(oh.dsad_loss(z, y, nom), loss.item(), rtol=1e-5)
        np.testing.assert_allclose(oh.dsad_grad(z, y, nom), zt.grad.numpy(), rtol=3e-5, atol=1e-9)
    np.testing.assert_allclose(oh.dsad_score(z), ref["DSADTrainer"].compute_anomaly_score(None, zt.detach(), None).numpy(),
                               rtol=1e-5, atol=1e-8)
    zt.grad = None
    loss = ref["DSVDDTrainer"].loss(None, zt, None, torch.from_numpy(c))
    loss.backward()
    np.testing.assert_allclose(oh.dsvdd_loss(z, c), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(oh.dsvdd_grad(z, c), zt.grad.numpy(), rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(oh.dsvdd_score(z, c),
                               ref["DSVDDTrainer"].compute_anomaly_score(None, zt.detach(), torch.from_numpy(c)).numpy(), rtol=1e-5)


@pytest.mark.parametrize("n", [2, 257])
def test_focal_live(ref, n):
    rng = np.random.default_rng(n + 3)
    x = (4 * rng.standard_normal((n, 1))).astype(np.float32)
    x[0, 0] = 30.0                                        # exp(-bce) below eps when the label disagrees: clamp branch
    y = rng.integers(0, 2, n)
    y[0] = 0
    xt = torch.from_numpy(x).requires_grad_(True)
    loss = ref["FocalTrainer"].loss(None, xt, torch.from_numpy(y), None)
    loss.backward()
    np.testing.assert_allclose(oh.focal_loss(x, y), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(oh.focal_grad(x, y), xt.grad.numpy().reshape(-1), rtol=3e-5, atol=1e-10)
    for nom in (0, 1):
        sc = ref["FocalTrainer"].compute_anomaly_score(None, xt.detach(), None, nominal_label=nom).numpy()
        np.testing.assert_allclose(oh.focal_score(x, nom), sc, rtol=3e-6, atol=1e-12)   # 1 - sigmoid: one fp32 ulp


@pytest.mark.parametrize("K", [2, 5, 30])
@pytest.mark.parametrize("mode", ["one_vs_rest", "leave_one_out"])
def test_clip_live(ref, K, mode):
    rng = np.random.default_rng(K)
    n, d = 33, 512
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = rng.standard_normal((K, d)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    y = rng.integers(0, 2, n)
    zt = torch.from_numpy(z).requires_grad_(True)
    sc = ref["ADClipTrainer"].compute_anomaly_score(_Self(mode), zt.detach(), torch.from_numpy(c)).numpy()
    np.testing.assert_allclose(oh.clip_score(z, c), sc, rtol=3e-4, atol=1e-30)
    for nom in (0, 1):
        loss = ref["ADClipTrainer"].loss(_Self(mode), zt, torch.from_numpy(y), torch.from_numpy(c), nominal_label=nom)
        zt.grad = None
        loss.backward()
        loo = mode == "leave_one_out"
        np.testing.assert_allclose(oh.clip_oe_loss(z, y, c, nom, loo), loss.item(), rtol=2e-5)
        np.testing.assert_allclose(oh.clip_oe_grad(z, y, c, nom, loo), zt.grad.numpy(), rtol=2e-3, atol=2e-7)


def test_vit_live_small(ref):
    """A 2-layer tower keeps this fast; full 12-layer parity is pinned by tests/golden/vit_*.npz."""
    for patch in (32, 16):
        sd = ovit.synth_state_dict(patch, seed=5, layers=2)
        m = ref["VisualTransformer"](224, patch, 768, 2, 12, 512).eval()
        m.load_state_dict({k[len("visual."):]: v for k, v in sd.items()})
        x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(9))
        with torch.no_grad():
            want = m(x)
        got = ovit.encode_image(sd, x)
        assert (want - got).abs().max().item() < 2e-5


def test_ref_fp16_emulation(ref):
    """oracle.vit.encode_image_ref_fp16 (the reference's GPU precision, model.py:371-392) against the live reference run
    in half on the CPU: individual roundings decorrelate the two (they differ by ~1e-3, like two noise draws), so the pin
    is statistical: both sit at the same distance from the fp32 answer (within 15 %)."""
    from oracle.make_golden import live_half_features
    for patch in (32, 16):
        sd = ovit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
        imgs = gi.vit_images(B=4, seed=77)
        f32 = ovit.encode_image(sd, imgs)
        live = torch.from_numpy(live_half_features(ref, patch, sd, imgs))
        emu = ovit.encode_image_ref_fp16(sd, imgs)
        d_live = ((live - f32).norm() / f32.norm()).item()
        d_emu = ((emu - f32).norm() / f32.norm()).item()
        assert 0.85 < d_emu / d_live < 1.15, (d_emu, d_live)
        assert 8e-4 < d_live < 2e-3


def test_text_live_small(ref):
    """CLIP.encode_text (model.py:339-352) on a 2-layer, small-vocabulary tower; the 12-layer tower is pinned by
    tests/golden/text.npz.  Rows cover the shortest prompt (<sot><eot>) and one that fills the context."""
    sd = otext.synth_text_state_dict(seed=5, layers=2, vocab=1000)
    m = ref["CLIP"](512, 224, 2, 768, 32, 77, 1000, 512, 8, 2).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("visual.") or k == "logit_scale" for k in missing)
    tok = otext.synth_tokens(6, seed=2, vocab=1000)
    assert tok[0].argmax() == 1 and tok[-1].argmax() == 76
    with torch.no_grad():
        want = m.encode_text(tok)
    got = otext.encode_text(sd, tok)
    assert (want - got).abs().max().item() < 2e-5
    # a repeated maximum: argmax takes the first (the reference indexes with text.argmax(dim=-1))
    tok2 = tok.clone()
    tok2[2, 40] = 999
    tok2[2, 50] = 999
    with torch.no_grad():
        want2 = m.encode_text(tok2)
    assert (want2 - otext.encode_text(sd, tok2)).abs().max().item() < 2e-5
