"""GPU parity of the fused loss/score heads (through the C ABI) vs the CPU oracle (oracle/heads.py) and the
golden fixtures generated from the live reference.  Tolerance: 1e-3 relative (north_star), tighter where
the arithmetic allows; written next to each assert."""
import os

import numpy as np
import pytest
import torch

from oracle import golden_inputs as gi
from oracle import heads as oh

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _np(t):
    return t.detach().float().cpu().numpy()


def _close(got, want, rtol=1e-3, atol=1e-7):
    np.testing.assert_allclose(np.asarray(got, np.float64), np.asarray(want, np.float64), rtol=rtol, atol=atol)


def test_hsc_golden(golden_dir):
    from eoe_b200 import ops
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    z, y = gi.hsc_inputs()
    for nom in (0, 1):
        zt = _t(z).requires_grad_(True)
        loss, scores = ops.hsc_loss(zt, _t(y), nominal_label=nom)
        loss.backward()
        _close(loss.item(), g[f"hsc_loss_nom{nom}"], rtol=1e-5)
        _close(_np(zt.grad), g[f"hsc_grad_nom{nom}"], rtol=1e-4, atol=1e-9)
        _close(_np(scores), g["hsc_score"], rtol=1e-4, atol=1e-7)
    _close(_np(ops.hsc_score(_t(z))), g["hsc_score"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("n,d", [(1, 4), (3, 8), (7, 100), (33, 128), (256, 256), (257, 512), (64, 1024),
                                 (5, 33), (19, 2048), (1000, 260)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_hsc_vs_oracle(n, d, dtype):
    from eoe_b200 import ops
    rng = np.random.default_rng(n * 1000 + d)
    z = (0.8 / np.sqrt(d) * rng.standard_normal((n, d))).astype(np.float32)
    zt = _t(z, dtype)
    zq = _np(zt)                                   # oracle sees the same (rounded) inputs
    y = rng.integers(0, 2, n)
    for nom in (0, 1):
        zz = zt.clone().requires_grad_(True)
        loss, scores = ops.hsc_loss(zz, _t(y), nominal_label=nom)
        (loss * 3.0).backward()                    # upstream gradient is honoured
        _close(loss.item(), oh.hsc_loss(zq, y, nom), rtol=1e-4)
        _close(_np(scores), oh.hsc_score(zq), rtol=2e-4, atol=1e-7)
        gtol = 1e-4 if dtype == torch.float32 else 1e-2   # grads are rounded to the feature dtype
        _close(_np(zz.grad), 3.0 * oh.hsc_grad(zq, y, nom), rtol=gtol, atol=1e-6 if dtype != torch.float32 else 1e-9)


def test_hsc_nan_propagates_and_zero_row():
    from eoe_b200 import ops
    z = torch.zeros(4, 256, device=DEV)
    z[1, 5] = float("nan")
    z[2] = 0.01
    loss, scores, grad = ops.hsc_fused(z, torch.tensor([0, 1, 1, 1], device=DEV))
    s = _np(scores)
    assert s[0] == 0.0 and np.isnan(s[1]) and np.isfinite(s[2])
    assert np.isnan(loss.item())
    assert float(grad[0].abs().sum()) == 0.0 and float(grad[3].abs().sum()) == 0.0


def test_bce_golden(golden_dir):
    from eoe_b200 import ops
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    x, y = gi.bce_inputs()
    xt = _t(x).requires_grad_(True)
    loss, scores = ops.bce_loss(xt, _t(y))
    loss.backward()
    _close(loss.item(), g["bce_loss"], rtol=1e-5)
    _close(_np(xt.grad), g["bce_grad"], rtol=1e-4, atol=1e-10)
    for nom in (0, 1):
        _close(_np(ops.bce_score(_t(x), nominal_label=nom)), g[f"bce_score_nom{nom}"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 255, 256, 1001, 65537])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_bce_vs_oracle(n, dtype):
    from eoe_b200 import ops
    rng = np.random.default_rng(n)
    x = (4 * rng.standard_normal((n, 1))).astype(np.float32)
    xt = _t(x, dtype)
    xq = _np(xt)
    y = rng.integers(0, 2, n)
    xx = xt.clone().requires_grad_(True)
    loss, scores = ops.bce_loss(xx, _t(y), nominal_label=1)
    loss.backward()
    assert xx.grad.shape == xx.shape
    _close(loss.item(), oh.bce_loss(xq, y), rtol=1e-4)
    _close(_np(scores), oh.bce_score(xq, 1), rtol=1e-4, atol=1e-7)
    gtol = 1e-4 if dtype == torch.float32 else 1e-2
    _close(_np(xx.grad).reshape(-1), oh.bce_grad(xq, y), rtol=gtol, atol=1e-9 if dtype == torch.float32 else 1e-6)
    # unaligned views take the scalar path
    if n > 3:
        l2, s2, g2 = ops.bce_fused(xt.reshape(-1)[1:], _t(y)[1:], 0)
        _close(l2.item(), oh.bce_loss(xq[1:], y[1:]), rtol=1e-4)
        _close(_np(s2), oh.bce_score(xq[1:], 0), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("K", [2, 10, 30])
def test_clip_golden(golden_dir, K):
    from eoe_b200 import ops
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    z, y, c = gi.clip_inputs(K)
    _close(_np(ops.clip_score(_t(z), _t(c))), g[f"clip_score_K{K}"], rtol=1e-3, atol=1e-30)
    for mode in ("one_vs_rest", "leave_one_out"):
        for nom in (0, 1):
            zt = _t(z).requires_grad_(True)
            loss = ops.clip_oe_loss(zt, _t(y), _t(c), nominal_label=nom, leave_one_out=(mode == "leave_one_out"))
            loss.backward()
            _close(loss.item(), g[f"clip_loss_K{K}_{mode}_nom{nom}"], rtol=1e-4)
            _close(_np(zt.grad), g[f"clip_grad_K{K}_{mode}_nom{nom}"], rtol=2e-3, atol=2e-6)


@pytest.mark.parametrize("n,d,K", [(1, 512, 2), (37, 512, 5), (300, 512, 30), (64, 768, 33), (50, 1024, 40), (9, 64, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_clip_vs_oracle(n, d, K, dtype):
    from eoe_b200 import ops
    rng = np.random.default_rng(n + d + K)
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = (rng.standard_normal((K, d)) * 1.7).astype(np.float32)      # non-unit: the score path must renormalise
    zt = _t(z, dtype)
    zq = _np(zt)
    _close(_np(ops.clip_score(zt, _t(c))), oh.clip_score(zq, c), rtol=1e-3, atol=1e-30)
    cu = (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)
    y = rng.integers(0, 2, n)
    for loo in (False, True):
        zz = zt.clone().requires_grad_(True)
        loss = ops.clip_oe_loss(zz, _t(y), _t(cu), nominal_label=0, leave_one_out=loo)
        loss.backward()
        _close(loss.item(), oh.clip_oe_loss(zq, y, cu, 0, loo), rtol=1e-4)
        gtol = 2e-3 if dtype == torch.float32 else 2e-2
        _close(_np(zz.grad), oh.clip_oe_grad(zq, y, cu, 0, loo), rtol=gtol, atol=1e-5)


@pytest.mark.parametrize("n,d,K", [(2048, 512, 2), (4099, 512, 10), (5000, 512, 30), (2050, 512, 32), (3001, 1024, 17),
                                   (2100, 128, 9), (2048, 256, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_clip_score_tensor_core_path_vs_oracle(n, d, K, dtype):
    """n >= 2048, d % 128 == 0, K <= 32: clip_score_mma_kernel (mma.sync, hi/lo split operands).  Same 1e-3 bar on the
    scores as the FP32-FMA kernel, including far-tail scores (atol 1e-30): the split must keep the logits to ~1e-4."""
    from eoe_b200 import ops
    rng = np.random.default_rng(n + d + K)
    z = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.01, 30.0, (n, 1)).astype(np.float32)
    c = (rng.standard_normal((K, d)) * 1.7).astype(np.float32)      # non-unit: the score path must renormalise
    z[: n // 3] += 0.2 * np.sqrt(d) * c[rng.integers(0, K, n // 3)]    # some prompts win clearly: scores down to ~1e-10
    zt = _t(z, dtype)
    zq = _np(zt)
    got = _np(ops.clip_score(zt, _t(c)))
    want = oh.clip_score(zq, c)
    _close(got, want, rtol=1e-3, atol=1e-30)
    # the small-n kernel (FP32 FMA) on a slice gives the same scores
    _close(_np(ops.clip_score(zt[:100].contiguous(), _t(c))), got[:100], rtol=1e-3, atol=1e-30)


@pytest.mark.parametrize("n,d,K", [(2048, 512, 2), (4099, 512, 10), (5000, 512, 30), (2050, 256, 32), (2100, 128, 17)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("loo", [False, True])
def test_clip_oe_loss_tensor_core_path_vs_oracle(n, d, K, dtype, loo):
    """clip_oe_loss_mma_kernel (forward and backward products on mma.sync): loss 1e-4, gradients at the FP32-FMA kernel's
    bars.  Leave-one-out rows whose two best nominal logits are closer than 1e-3 are excluded from the gradient check: the
    argmax target (clip.py:95) may legitimately differ between two roundings of the same logits."""
    from eoe_b200 import ops
    rng = np.random.default_rng(n + d + K + loo)
    z = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.05, 20.0, (n, 1)).astype(np.float32)
    cu = rng.standard_normal((K, d)).astype(np.float32)
    cu = (cu / np.linalg.norm(cu, axis=1, keepdims=True)).astype(np.float32)
    z[: n // 3] += 0.1 * np.sqrt(d) * cu[rng.integers(0, K, n // 3)] * np.linalg.norm(z[: n // 3], axis=1, keepdims=True) / np.sqrt(d)
    y = rng.integers(0, 2, n)
    y[5] = 2                                             # label outside {0,1}: contributes 0 (clip.py:90-92)
    zt = _t(z, dtype)
    zq = _np(zt)
    for nom in (0, 1):
        zz = zt.clone().requires_grad_(True)
        loss = ops.clip_oe_loss(zz, _t(y), _t(cu), nominal_label=nom, leave_one_out=loo)
        loss.backward()
        _close(loss.item(), oh.clip_oe_loss(zq, y, cu, nom, loo), rtol=1e-4)
        got, want = _np(zz.grad), oh.clip_oe_grad(zq, y, cu, nom, loo)
        keep = np.ones(n, bool)
        if loo and K > 2:
            lg = np.sort(oh.clip_logits(zq, cu, False, 100.0)[:, : K - 1], axis=1)
            keep = (lg[:, -1] - lg[:, -2]) > 1e-3
            assert keep.mean() > 0.99
        assert np.all(got[5] == 0)
        gtol = 2e-3 if dtype == torch.float32 else 2e-2
        _close(got[keep], want[keep], rtol=gtol, atol=1e-5 / n * 128)
        # loss only (no gradient buffer) takes the same forward
        l2, g2 = ops.clip_oe_fused(zt, _t(y), _t(cu), nom, loo, want_grad=False)
        assert g2 is None and l2.item() == loss.item()


@pytest.mark.parametrize("n,d,K", [(300, 512, 47), (128, 512, 100), (4500, 512, 200), (33, 512, 256), (70, 768, 65)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_clip_large_prompt_sets(n, d, K, dtype):
    """leave_one_out builds one prompt per class (+ the anomaly prompt): 47 (dtd), 100 (cifar100), 200 (cub) rows of
    `center` (clip.py:53-54).  Past 64 prompts the text rows are read from global memory instead of the shared-memory
    tile and the loss keeps 8 logit slots per lane; same bars as the small-K kernels."""
    from eoe_b200 import ops
    rng = np.random.default_rng(n + d + K)
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = (rng.standard_normal((K, d)) * 1.7).astype(np.float32)
    z[: n // 2] += 0.2 * np.sqrt(d) * c[rng.integers(0, K, n // 2)] / 1.7
    zt = _t(z, dtype)
    zq = _np(zt)
    _close(_np(ops.clip_score(zt, _t(c))), oh.clip_score(zq, c), rtol=1e-3, atol=1e-30)
    cu = (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)
    y = rng.integers(0, 2, n)
    for loo in (False, True):
        for nom in (0, 1):
            zz = zt.clone().requires_grad_(True)
            loss = ops.clip_oe_loss(zz, _t(y), _t(cu), nominal_label=nom, leave_one_out=loo)
            loss.backward()
            _close(loss.item(), oh.clip_oe_loss(zq, y, cu, nom, loo), rtol=1e-4)
            keep = np.ones(n, bool)
            if loo:
                lg = np.sort(oh.clip_logits(zq, cu, False, 100.0)[:, : K - 1], axis=1)
                keep = (lg[:, -1] - lg[:, -2]) > 1e-3
            gtol = 2e-3 if dtype == torch.float32 else 2e-2
            _close(_np(zz.grad)[keep], oh.clip_oe_grad(zq, y, cu, nom, loo)[keep], rtol=gtol, atol=1e-5)


@pytest.mark.parametrize("n,d,K", [(16384, 512, 30), (20001, 512, 10), (16500, 512, 32), (17000, 256, 2), (16385, 64, 1),
                                   (40000, 384, 17)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_clip_score_tcgen05_path_vs_oracle(n, d, K, dtype):
    """16-bit rows, n >= 16 384, d % 64 == 0, d <= 512, K <= 32: clip_score_tc_kernel (tcgen05 / TMEM, rows brought by TMA and
    used as UMMA operands as they are, text hi ; lo as one N = 64 operand).  Same bars as the warp-level kernel: 1e-3 on the
    scores including far-tail ones, ragged last tile, and agreement with the warp-level kernel on a slice below the
    dispatch threshold."""
    from eoe_b200 import ops
    rng = np.random.default_rng(n + d + K)
    z = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.01, 30.0, (n, 1)).astype(np.float32)
    c = (rng.standard_normal((K, d)) * 1.7).astype(np.float32)
    z[: n // 3] += 0.2 * np.sqrt(d) * c[rng.integers(0, K, n // 3)]
    zt = _t(z, dtype)
    zq = _np(zt)
    got = _np(ops.clip_score(zt, _t(c)))
    _close(got, oh.clip_score(zq, c), rtol=1e-3, atol=1e-30)
    if d % 128 == 0:                                        # the mma.sync kernel on the first 4 000 rows
        _close(_np(ops.clip_score(zt[:4000].contiguous(), _t(c))), got[:4000], rtol=1e-3, atol=1e-30)
    else:
        _close(_np(ops.clip_score(zt[:100].contiguous(), _t(c))), got[:100], rtol=1e-3, atol=1e-30)


@pytest.mark.parametrize("n,d,K", [(16384, 512, 30), (20001, 512, 10), (16500, 512, 32), (17000, 256, 2), (16385, 64, 1),
                                   (33000, 384, 17)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("loo", [False, True])
def test_clip_oe_loss_tcgen05_path_vs_oracle(n, d, K, dtype, loo):
    """16-bit rows, n >= 16 384: clip_oe_loss_tc_kernel (forward logits and G @ C on tcgen05, G handed over in TMEM, dz written
    in place over the z chunks and TMA-stored).  Same bars as the warp-level kernel: loss 1e-4, gradients 2e-2 of the 16-bit
    outputs (leave-one-out rows whose two best nominal logits nearly tie are excluded: their arg-max target legitimately
    depends on the last bit), labels outside {0, 1} contribute nothing, ragged last tile."""
    from eoe_b200 import ops
    if loo and K == 1:
        pytest.skip("leave-one-out needs at least one nominal prompt beside the anomaly prompt")
    rng = np.random.default_rng(n + d + K + int(loo))
    z = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.5, 4.0, (n, 1)).astype(np.float32)
    c = rng.standard_normal((K, d)).astype(np.float32)
    cu = c / np.linalg.norm(c, axis=1, keepdims=True)
    y = rng.integers(0, 2, n).astype(np.int64)
    y[::17] = -1                                            # unlabeled rows: no loss, zero gradient
    for nom in (0, 1):
        zt = _t(z, dtype).requires_grad_(True)
        zq = _np(zt.detach())
        loss = ops.clip_oe_loss(zt, _t(y), _t(cu), nom, leave_one_out=loo)
        loss.backward()
        np.testing.assert_allclose(loss.item(), oh.clip_oe_loss(zq, y, cu, nom, loo), rtol=1e-4)
        g = _np(zt.grad)
        want = oh.clip_oe_grad(zq, y, cu, nom, loo)
        keep = np.ones(n, bool)
        if loo and K > 2:
            lg = np.sort(oh.clip_logits(zq, cu, False, 100.0)[:, : K - 1], axis=1)
            keep = (lg[:, -1] - lg[:, -2]) > 1e-3
            assert keep.mean() > 0.99
        assert np.all(g[y == -1] == 0)
        _close(g[keep], want[keep], rtol=2e-2, atol=1e-5 / n * 128)


def test_clip_tcgen05_heads_many_tiles_per_sm_are_deterministic():
    """Steady state of the persistent kernels (the parity cases above give an SM at most two tiles): 100 000 rows = 5-6 tiles
    per SM against the oracle, then 2^20 rows = 55 tiles per SM launched 30 times -- every launch returns bit-identical
    loss, gradients and scores (fixed reduction order, no atomics), i.e. the producer / MMA / owner handshakes carry no race."""
    from eoe_b200 import ops
    rng = np.random.default_rng(11)
    n, d, K = 100000, 512, 30
    z = rng.standard_normal((n, d)).astype(np.float32)
    cu = rng.standard_normal((K, d)).astype(np.float32)
    cu /= np.linalg.norm(cu, axis=1, keepdims=True)
    y = rng.integers(0, 2, n).astype(np.int64)
    zt = _t(z, torch.bfloat16).requires_grad_(True)
    zq = _np(zt.detach())
    loss = ops.clip_oe_loss(zt, _t(y), _t(cu), 0)
    loss.backward()
    np.testing.assert_allclose(loss.item(), oh.clip_oe_loss(zq, y, cu, 0, False), rtol=1e-4)
    _close(_np(zt.grad), oh.clip_oe_grad(zq, y, cu, 0, False), rtol=2e-2, atol=1e-5 / n * 128)
    _close(_np(ops.clip_score(zt.detach(), _t(cu))), oh.clip_score(zq, cu), rtol=1e-3, atol=1e-30)
    n = 1 << 20
    g = torch.Generator(device=DEV).manual_seed(3)
    zb = torch.randn(n, d, device=DEV, generator=g).to(torch.float16)
    yb = torch.randint(0, 2, (n,), device=DEV, generator=g)
    ct = _t(cu)
    l0, g0 = ops.clip_oe_fused(zb, yb, ct, 0, True)
    s0 = ops.clip_score(zb, ct)
    l0, g0, s0 = l0.clone(), g0.clone(), s0.clone()
    assert torch.isfinite(g0.float()).all() and torch.isfinite(s0).all()
    for _ in range(30):
        l1, g1 = ops.clip_oe_fused(zb, yb, ct, 0, True)
        s1 = ops.clip_score(zb, ct)
        assert l1.item() == l0.item() and torch.equal(g1, g0) and torch.equal(s1, s0)


def test_clip_score_tcgen05_path_special_rows_and_determinism():
    from eoe_b200 import ops
    rng = np.random.default_rng(5)
    n, d, K = 16384 + 77, 512, 30
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = rng.standard_normal((K, d)).astype(np.float32)
    z[7, 100] = np.nan
    z[8, 5] = np.inf
    z[24] = 0.0
    z[n - 1, 511] = np.nan
    zt = _t(z, torch.float16)
    got = ops.clip_score(zt, _t(c))
    assert torch.equal(got, ops.clip_score(zt, _t(c)).clone()) or torch.equal(torch.nan_to_num(got), torch.nan_to_num(ops.clip_score(zt, _t(c))))
    got = _np(got)
    bad = np.zeros(n, bool)
    bad[[7, 8, 24, n - 1]] = True
    assert np.isnan(got[bad]).all() and np.isfinite(got[~bad]).all()
    _close(got[~bad], oh.clip_score(_np(zt)[~bad], c), rtol=1e-3, atol=1e-30)


def test_clip_score_tensor_core_path_special_rows():
    """NaN / Inf / zero rows give NaN scores (z / ||z|| in the reference, clip.py:70), neighbours are untouched."""
    from eoe_b200 import ops
    rng = np.random.default_rng(3)
    n, d, K = 2048 + 5, 512, 10
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = rng.standard_normal((K, d)).astype(np.float32)
    z[7, 100] = np.nan
    z[8, 5] = np.inf
    z[24] = 0.0
    z[n - 1, 511] = np.nan
    got = _np(ops.clip_score(_t(z), _t(c)))
    bad = np.zeros(n, bool)
    bad[[7, 8, 24, n - 1]] = True
    assert np.isnan(got[bad]).all() and np.isfinite(got[~bad]).all()
    _close(got[~bad], oh.clip_score(z[~bad], c), rtol=1e-3, atol=1e-30)


def test_bad_arguments_raise():
    from eoe_b200 import _lib, ops
    with pytest.raises(_lib.EoeError):
        ops.hsc_score(torch.zeros(4, 8))                      # CPU tensor: no fallback
    with pytest.raises(_lib.EoeError):
        ops.clip_score(torch.zeros(4, 510, device=DEV), torch.zeros(3, 510, device=DEV))   # d % 4 != 0
    with pytest.raises(_lib.EoeError):
        ops.clip_score(torch.zeros(4, 512, device=DEV), torch.zeros(257, 512, device=DEV))  # K > 256 prompts


def test_heads_large_property():
    """BASELINE-size rows (2^20 x 256): size-independent properties instead of a CPU oracle pass:
    score/loss of a tiled input equal the small-case values; grad is coef*z (linearity in z per row)."""
    from eoe_b200 import ops
    rng = np.random.default_rng(0)
    base = (0.05 * rng.standard_normal((1024, 256))).astype(np.float32)
    yb = rng.integers(0, 2, 1024)
    z = _t(base).repeat(1024, 1)
    y = _t(yb).repeat(1024)
    loss, scores, grad = ops.hsc_fused(z, y, 0)
    _close(loss.item(), oh.hsc_loss(base, yb, 0), rtol=1e-4)
    sc = _np(scores).reshape(1024, 1024)
    assert np.array_equal(sc[0], sc[-1]) and np.array_equal(sc[0], sc[511])
    _close(sc[0], oh.hsc_score(base), rtol=2e-4, atol=1e-7)
    g = _np(grad[:1024]) * 1024.0
    _close(g, oh.hsc_grad(base, yb, 0), rtol=1e-4, atol=1e-9)
    assert torch.equal(grad[:1024], grad[-1024:])


# ------------------------------------------------------------------------------------------------ DSAD / DSVDD / focal
def test_dsad_dsvdd_focal_golden(golden_dir):
    """The reference's own dsad.py / dsvdd.py / focal.py hooks (fixtures from oracle/make_golden.py)."""
    from eoe_b200 import ops
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    z, y = gi.hsc_inputs()
    for nom in (0, 1):
        zt = _t(z).requires_grad_(True)
        loss, scores = ops.dsad_loss(zt, _t(y), nominal_label=nom)
        loss.backward()
        _close(loss.item(), g[f"dsad_loss_nom{nom}"], rtol=1e-5)
        # the zero row of an anomalous sample has loss 1e9 and gradient 0 * 1e18: both sides give 0
        _close(_np(zt.grad), g[f"dsad_grad_nom{nom}"], rtol=1e-4, atol=1e-9)
        _close(_np(scores), g["dsad_score"], rtol=1e-4, atol=1e-7)
    zd, cd = gi.dsvdd_inputs()
    zt = _t(zd).requires_grad_(True)
    loss, scores = ops.dsvdd_loss(zt, _t(cd))
    loss.backward()
    _close(loss.item(), g["dsvdd_loss"], rtol=1e-5)
    _close(_np(scores), g["dsvdd_score"], rtol=1e-5, atol=1e-7)
    _close(_np(zt.grad), g["dsvdd_grad"], rtol=1e-5, atol=1e-10)
    _close(_np(ops.dsvdd_score(_t(zd), _t(cd))), g["dsvdd_score"], rtol=1e-5, atol=1e-7)
    x, yb = gi.bce_inputs()
    xt = _t(x).requires_grad_(True)
    loss, scores = ops.focal_loss(xt, _t(yb))
    loss.backward()
    _close(loss.item(), g["focal_loss"], rtol=1e-5)
    _close(_np(xt.grad), g["focal_grad"], rtol=1e-4, atol=1e-10)
    _close(_np(scores), g["focal_score_nom0"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n,d", [(1, 4), (7, 100), (256, 256), (257, 512), (5, 33), (19, 2048), (1000, 260)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_dsad_dsvdd_vs_oracle(n, d, dtype):
    from eoe_b200 import ops
    rng = np.random.default_rng(n * 1000 + d + 1)
    z = (0.8 / np.sqrt(d) * rng.standard_normal((n, d))).astype(np.float32)
    c = (0.5 / np.sqrt(d) * rng.standard_normal((1, d))).astype(np.float32)
    zt = _t(z, dtype)
    zq = _np(zt)
    y = rng.integers(0, 2, n)
    gtol = 1e-4 if dtype == torch.float32 else 1e-2       # grads are rounded to the feature dtype
    gatol = 1e-9 if dtype == torch.float32 else 1e-5
    for nom in (0, 1):
        zz = zt.clone().requires_grad_(True)
        loss, scores = ops.dsad_loss(zz, _t(y), nominal_label=nom)
        (loss * 2.0).backward()
        _close(loss.item(), oh.dsad_loss(zq, y, nom), rtol=1e-4)
        _close(_np(scores), oh.dsad_score(zq), rtol=2e-4, atol=1e-7)
        _close(_np(zz.grad), 2.0 * oh.dsad_grad(zq, y, nom), rtol=gtol, atol=gatol * max(1.0, float(np.abs(oh.dsad_grad(zq, y, nom)).max())))
    zz = zt.clone().requires_grad_(True)
    loss, scores = ops.dsvdd_loss(zz, _t(c))
    loss.backward()
    _close(loss.item(), oh.dsvdd_loss(zq, c), rtol=1e-4)
    _close(_np(scores), oh.dsvdd_score(zq, c), rtol=1e-4, atol=1e-7)
    _close(_np(zz.grad), oh.dsvdd_grad(zq, c), rtol=gtol, atol=gatol)
    _close(_np(ops.dsvdd_score(zt, _t(c))), oh.dsvdd_score(zq, c), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("n", [1, 2, 5, 255, 1001, 65537])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_focal_vs_oracle(n, dtype):
    from eoe_b200 import ops
    rng = np.random.default_rng(n + 17)
    x = (4 * rng.standard_normal((n, 1))).astype(np.float32)
    x[0, 0] = 30.0                                         # pt below eps for label 0: the clamp branch (gradient cut)
    xt = _t(x, dtype)
    xq = _np(xt)
    y = rng.integers(0, 2, n)
    y[0] = 0
    xx = xt.clone().requires_grad_(True)
    loss, scores = ops.focal_loss(xx, _t(y), nominal_label=1)
    loss.backward()
    _close(loss.item(), oh.focal_loss(xq, y), rtol=1e-4)
    _close(_np(scores), oh.focal_score(xq, 1), rtol=1e-4, atol=1e-7)
    gtol = 2e-4 if dtype == torch.float32 else 1e-2
    _close(_np(xx.grad).reshape(-1), oh.focal_grad(xq, y), rtol=gtol, atol=1e-9 if dtype == torch.float32 else 1e-6)


def test_new_heads_nan_propagates():
    from eoe_b200 import ops
    z = torch.full((4, 64), 0.1, device=DEV)
    z[1, 3] = float("nan")
    y = torch.tensor([0, 1, 0, 1], device=DEV)
    l, s, _ = ops.dsad_fused(z, y)
    assert np.isnan(l.item()) and np.isnan(_np(s)[1]) and np.isfinite(_np(s)[0])
    l, s, _ = ops.dsvdd_fused(z, torch.zeros(64, device=DEV))
    assert np.isnan(l.item()) and np.isnan(_np(s)[1]) and np.isfinite(_np(s)[2])
    x = torch.tensor([0.5, float("nan"), -1.0], device=DEV)
    l, s, g = ops.focal_fused(x, torch.tensor([1, 0, 1], device=DEV))
    assert np.isnan(l.item()) and np.isnan(_np(s)[1]) and np.isnan(_np(g)[1]) and np.isfinite(_np(g)[0])
