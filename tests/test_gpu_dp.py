"""GPU, >= 2 devices (skipped otherwise): data-parallel training through the re-hosted trainer over NCCL equals the
single-process run on the global batch, including the BatchNorm caveat SURVEY section 7 raised:

  * a model without batch statistics: DP == global batch to ~1e-5 (gradient buckets average equal shards; only the
    summation order differs)
  * BatchNorm under plain DP normalises each rank's 128 || 128 shard with its OWN statistics: the run DIFFERS from the
    global-batch run (delta reported, asserted to be visible) -- this is the reference's semantics only at world size 1
  * the same model converted with torch.nn.SyncBatchNorm: statistics are all-reduced, DP == global batch again (1e-4)

Run on the GPU box with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(kind):
    torch.manual_seed(0)
    nn = torch.nn
    norm = (lambda c: nn.BatchNorm2d(c, eps=1e-4, affine=False)) if kind != "plain" else (lambda c: nn.Identity())
    m = nn.Sequential(nn.Conv2d(3, 16, 5, padding=2, bias=False), norm(16), nn.LeakyReLU(), nn.MaxPool2d(2),
                      nn.Conv2d(16, 32, 5, padding=2, bias=False), norm(32), nn.LeakyReLU(), nn.MaxPool2d(2),
                      nn.Flatten(), nn.Linear(32 * 8 * 8, 64, bias=False))
    return m


def _batches(ws, n_batches=3, per=64):
    """global batches of ws * per rows; rank r owns rows [r * per, (r + 1) * per), each shard = per/2 normal || per/2 OE"""
    g = torch.Generator().manual_seed(3)
    out = []
    for _ in range(n_batches):
        imgs = torch.randn(ws * per, 3, 32, 32, generator=g)
        lbls = ((torch.arange(ws * per) % per) >= per // 2).long()
        imgs[lbls == 1] += 0.5
        out.append((imgs, lbls))
    return out


def _worker(rank, ws, port, kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    from eoe_b200 import dist as edist
    from eoe_b200.training import TRAINER
    edist.init_from_env()
    dev = torch.device("cuda", rank)
    model = _model(kind)
    if kind == "syncbn":
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    per = 64
    loader = [(x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per], None) for x, y in _batches(ws)]
    tr = TRAINER["hsc"](model, epochs=2, lr=1e-2, device=dev, data_parallel=True, sgd=True)
    model, roc, losses = tr.train_cls(model, loader, nominal_label=0)
    q.put((rank, [p.detach().cpu().numpy() for p in model.parameters()], losses, roc.auc))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def _single(kind, ws):
    from eoe_b200.training import TRAINER
    model = _model(kind)
    loader = [(x, y, None) for x, y in _batches(ws)]
    tr = TRAINER["hsc"](model, epochs=2, lr=1e-2, device="cuda:0", sgd=True)
    model, roc, losses = tr.train_cls(model, loader, nominal_label=0)
    return [p.detach().cpu().numpy() for p in model.parameters()], losses, roc.auc


@pytest.mark.parametrize("kind", ["plain", "bn", "syncbn"])
def test_dp_training_equals_global_batch_nccl_world2(kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, ws, port, kind, q)) for r in range(ws)]
    [p.start() for p in ps]
    res = {r: (w, l, a) for r, w, l, a in (q.get(timeout=300) for _ in ps)}
    [p.join(60) for p in ps]
    want_w, want_l, want_auc = _single(kind, ws)
    # every rank ends with the same weights, the same epoch AUC (scores are all-gathered) ...
    for a, b in zip(res[0][0], res[1][0]):
        np.testing.assert_array_equal(a, b)
    assert res[0][2] == res[1][2]
    dw = max(float(np.abs(a - b).max()) for a, b in zip(res[0][0], want_w))
    # the per-rank epoch loss is the mean over the rank's shard; the global loss is the mean of the two
    dl = max(abs((l0 + l1) / 2 - w) / abs(w) for l0, l1, w in zip(res[0][1], res[1][1], want_l))
    print("DP_VS_GLOBAL", dict(kind=kind, max_abs_weight_diff=dw, max_rel_loss_diff=dl, auc_dp=res[0][2], auc_global=want_auc))
    if kind == "plain":
        # measured on 2 x B200: 1.1e-5 on the weights after 6 SGD steps (lr 1e-2, momentum 0.9), 3.6e-5 on the epoch losses:
        # the all-reduce adds the two shard gradients in another order than one kernel sums the global batch
        assert dw < 5e-5 and dl < 1e-4
        assert abs(res[0][2] - want_auc) < 1e-6
    elif kind == "syncbn":
        assert dw < 1e-4 and dl < 1e-4
    else:
        assert dw > 1e-5                      # per-rank statistics: NOT the global-batch run (documented caveat) ...
        assert dl < 0.2                       # ... but the same optimisation problem up to the statistics' noise
