#!/usr/bin/env python
"""Multi-GPU (NCCL over NVLink) check of SURVEY 8(e), launched with torchrun on N GPUs:
  1. data-parallel training (HSC and BCE heads, GradBuckets all-reduce overlapped with backward): after 6 steps the
     weights and per-step losses equal a single-process run on the concatenated (global) batch;
  2. sharded evaluation: rank r scores its contiguous block, all_gather(scores, labels), every rank computes the global
     ROC-AUC on the device -- identical on all ranks and equal to the single-process value (bit-exact).
Prints one JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402
from eoe_b200 import dist as edist, metrics, ops  # noqa: E402


def model_for(objective, d_in=64):
    torch.manual_seed(0)
    d_out = 1 if objective == "bce" else 128
    return torch.nn.Sequential(torch.nn.Linear(d_in, 256), torch.nn.LeakyReLU(), torch.nn.Linear(256, d_out))


def loss_of(objective, f, y):
    return (ops.hsc_loss(f, y, 0) if objective == "hsc" else ops.bce_loss(f, y, 0))[0]


def main():
    rank, local, ws = edist.init_from_env()
    dev = torch.device("cuda", local)
    out = {"world_size": ws}
    per = 256                                                     # 128 normal || 128 OE per rank (bases.py:591-597)
    g = torch.Generator().manual_seed(7)
    steps = 6
    X = torch.randn(steps, ws * per, 64, generator=g)
    Y = torch.cat([torch.zeros(per // 2, dtype=torch.long), torch.ones(per // 2, dtype=torch.long)]).repeat(ws)
    for objective in ("hsc", "bce"):
        m_dp = model_for(objective).to(dev)
        m_ref = copy.deepcopy(m_dp)
        opt_dp = torch.optim.Adam(m_dp.parameters(), lr=1e-3)
        opt_ref = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
        buckets = edist.GradBuckets(list(m_dp.parameters()), bucket_bytes=64 << 10)     # several buckets -> real overlap
        max_dloss = 0.0
        for s in range(steps):
            xs, ys = X[s, rank * per:(rank + 1) * per].to(dev), Y[rank * per:(rank + 1) * per].to(dev)
            buckets.zero_grad()
            loss = loss_of(objective, m_dp(xs), ys)
            loss.backward()
            buckets.finish()
            opt_dp.step()
            # single-process arm on the global batch
            opt_ref.zero_grad()
            loss_ref = loss_of(objective, m_ref(X[s].to(dev)), Y.to(dev))
            loss_ref.backward()
            opt_ref.step()
            lg = loss.detach().clone()
            tdist.all_reduce(lg)
            max_dloss = max(max_dloss, abs(float(lg) / ws - float(loss_ref.detach())) / abs(float(loss_ref.detach())))
        dw = max(float((p - q).abs().max()) for p, q in zip(m_dp.parameters(), m_ref.parameters()))
        out[objective] = {"max_rel_loss_diff": max_dloss, "max_abs_weight_diff": dw}
        assert max_dloss < 1e-4 and dw < 1e-4, out
    # sharded evaluation
    n = 1_000_003
    gs = torch.Generator().manual_seed(3)
    scores = (1 - torch.exp(-torch.randn(n, generator=gs).abs())).to(dev)
    labels = (torch.rand(n, generator=gs) < 0.4).long().to(dev)
    lo, hi = edist.shard_range(n, rank, ws)
    s_all, l_all = edist.all_gather_rows(scores[lo:hi]), edist.all_gather_rows(labels[lo:hi])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); tdist.barrier()
    ws_auc = metrics.AucWorkspace().ensure(n, dev)
    for _ in range(2):
        a.record()
        s_all, l_all = edist.all_gather_rows(scores[lo:hi], total=n), edist.all_gather_rows(labels[lo:hi], total=n)
        auc = metrics.roc_auc_device(s_all, l_all, workspace=ws_auc)[0][0]
        b.record(); torch.cuda.synchronize()
    auc_single = metrics.roc_auc_device(scores, labels)[0][0]
    gathered = [torch.zeros_like(auc) for _ in range(ws)]
    tdist.all_gather(gathered, auc)
    out["eval"] = {"n_scores": n, "auc": float(auc), "bit_exact_vs_single_process": bool(auc == auc_single),
                   "identical_on_all_ranks": bool(all(bool(x == auc) for x in gathered)),
                   "gather_plus_auc_ms": a.elapsed_time(b)}
    assert out["eval"]["bit_exact_vs_single_process"] and out["eval"]["identical_on_all_ranks"], out
    if rank == 0:
        print(json.dumps(out), flush=True)
    tdist.barrier()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
