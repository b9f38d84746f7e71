"""CPU: host-side logic -- shard math, the N>1 collectives on gloo (world_size 2), gradient buckets, trainer plumbing."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from eoe_b200 import dist as edist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for n in [0, 1, 7, 8, 9, 3000, 10_000, 1_000_000]:
        for w in [1, 2, 3, 4, 8]:
            spans = [edist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(lo <= hi for lo, hi in spans)
            assert max(hi - lo for lo, hi in spans) == (n + w - 1) // w if n else True


def _worker_gather(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    edist.init_from_env(backend="gloo")
    n = 1001
    full_s = torch.arange(n, dtype=torch.float32) * 0.5
    full_l = (torch.arange(n) % 3 == 0).long()
    lo, hi = edist.shard_range(n, rank, ws)
    s = edist.all_gather_rows(full_s[lo:hi])
    l = edist.all_gather_rows(full_l[lo:hi])
    ok = torch.equal(s, full_s) and torch.equal(l, full_l)
    # ragged + empty shard
    t = edist.all_gather_rows(torch.full((rank * 3,), float(rank)))
    ok = ok and torch.equal(t, torch.cat([torch.full((r * 3,), float(r)) for r in range(ws)]))
    q.put((rank, ok))
    torch.distributed.destroy_process_group()


def test_all_gather_rows_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_gather, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(30) for p in ps]
    assert all(ok for _, ok in res)


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8))


def _hsc_loss_torch(z, y):      # plain torch statement of hsc.py:17-21, only to drive the gradient plumbing on CPU
    d = torch.sqrt(torch.norm(z, p=2, dim=1) ** 2 + 1) - 1
    s = 1 - torch.exp(-d)
    return torch.where(y == 0, d, -torch.log(s + 1e-9)).mean()


def _worker_dp(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    edist.init_from_env(backend="gloo")
    model = _make_model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 16, generator=g)
    y = (torch.rand(64, generator=g) < 0.5).long()
    lo, hi = edist.shard_range(64, rank, ws)
    buckets = edist.GradBuckets(model.parameters(), bucket_bytes=1024)     # several small buckets
    for _ in range(2):                                                      # re-arming works
        buckets.zero_grad()
        _hsc_loss_torch(model(x[lo:hi]), y[lo:hi]).backward()
        buckets.finish()
    grads = [p.grad.clone().numpy() for p in model.parameters()]     # numpy: pickled by value (a torch tensor would be
    q.put((rank, grads))                                              # passed as an fd that dies with this process)
    torch.distributed.destroy_process_group()


def test_grad_buckets_average_equals_full_batch_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_dp, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    model = _make_model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 16, generator=g)
    y = (torch.rand(64, generator=g) < 0.5).long()
    _hsc_loss_torch(model(x), y).backward()       # mean over the global batch == average of equal-shard means
    for r in (0, 1):
        for got, p in zip(res[r], model.parameters()):
            torch.testing.assert_close(torch.from_numpy(got), p.grad, rtol=1e-5, atol=1e-7)


def test_trainer_registry_and_hook_signatures():
    import inspect
    from eoe_b200.training import TRAINER, ADTrainer
    assert set(TRAINER) == {"hsc", "bce", "clip", "dsvdd", "dsad", "focal"}      # training/__init__.py:8-11
    for cls in TRAINER.values():
        assert issubclass(cls, ADTrainer)
        for hook, first in (("prepare_metric", ["self", "cstr", "loader", "model", "seed"]),
                            ("compute_anomaly_score", ["self"]), ("loss", ["self"])):
            params = list(inspect.signature(getattr(cls, hook)).parameters)
            assert params[:len(first)] == first
            assert "kwargs" in params
    t = TRAINER["hsc"](None, device="cpu")
    assert t.prepare_metric("x", None, None, 0) is None
    with pytest.raises(NotImplementedError):
        TRAINER["hsc"](None, ad_mode="fifty_fifty", device="cpu")


def test_clip_prompts_follow_reference():
    from eoe_b200.training import ADClipTrainer
    seen = {}

    def enc(texts):
        seen["t"] = list(texts)
        return torch.eye(len(texts), 512) * 3.0

    t = ADClipTrainer(None, device="cpu", text_encoder=enc, class_names=["cat", "dog", "ship"], ad_mode="leave_one_out")
    c = t.prepare_metric("dog", None, None, 0)
    assert seen["t"] == ["a photo of a cat", "a photo of a ship", "a photo of something"]
    assert torch.allclose(c.norm(dim=-1), torch.ones(3))
    t2 = ADClipTrainer(None, device="cpu", text_encoder=enc, anom_tkn_ptn="a photo of something that is not a {}")
    t2.prepare_metric("dog", None, None, 0)
    assert seen["t"] == ["a photo of a dog", "a photo of something that is not a dog"]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port arm the driver runs beside ours): one JSON line with the contract's keys;
    under a multi-rank launch only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--ref-images", "2", "--patch", "32", "--prompts", "10"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "clip_zero_shot_ad_images_per_s" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "clip_vitb32_zero_shot_ad_224px_10prompts"
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1"})
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]
