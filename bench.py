#!/usr/bin/env python
"""Headline benchmark: CLIP ViT-B/16 zero-shot anomaly-detection scoring (BASELINE.json configs[2]) in images/s.

One "step" = one batch of 224x224 images per GPU through the hot path: patchify -> ViT-B/16 encoder (tcgen05 GEMMs,
attention, LayerNorm) -> fused cosine-softmax anomaly score.  The timed region covers K steps, the score all-gather
(N > 1) and the global device ROC-AUC over every score produced.  See DESIGN.md "Measurement".

    python bench.py --gpus 1 --steps 8 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference          # the reference algorithm's CPU port on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "clip_zero_shot_ad_images_per_s"
UNIT = "images/s"
GFLOP_PER_IMG = {16: 35.12690688, 32: 8.81762304}      # 2*MAC over GEMMs + QK^T + PV


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi sampled every 50 ms in the background; only samples whose timestamp falls inside the timed region
    (marked with `begin()` / `end()`) are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/eoe_clocks_{os.getpid()}.csv"
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        if not inside and rows and self.t0 is not None:       # region shorter than the sampling period: nearest sample
            mid = self.t0 + (self.t1 - self.t0) / 2
            inside = [min(rows, key=lambda r: abs((r[0] - mid).total_seconds()))]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out.update(sm_mhz=statistics.median(r[1] for r in inside), sm_max_mhz=max(r[2] for r in inside),
                       reasons=sorted(reasons), samples=len(inside), power_w_max=max(r[3] for r in inside))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def cpu_port_images_per_s(patch, K, n_img, threads):
    """The reference algorithm on the host cores: oracle port of VisualTransformer.forward + CLIP score head (fp32)."""
    from oracle import heads as oh
    from oracle import vit as ovit
    torch.set_num_threads(threads)
    sd = ovit.synth_state_dict(patch, seed=0)
    g = torch.Generator().manual_seed(1)
    imgs = torch.randn(n_img, 3, 224, 224, generator=g)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1).numpy()
    chunk = 32
    ovit.encode_image(sd, imgs[:min(4, n_img)])          # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for s in range(0, n_img, chunk):
        f = ovit.encode_image(sd, imgs[s:s + chunk])
        oh.clip_score(f.numpy(), text)
    dt = time.perf_counter() - t0
    return n_img / dt, dt


_STDOUT_FD = None


def emit(line):
    """Print the result line on the real stdout (see main())."""
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _STDOUT_FD is not None:
        os.dup2(2, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_img = args.ref_images
    t_all, n_all = 0.0, 0
    for _ in range(args.warmup):
        cpu_port_images_per_s(args.patch, args.prompts, min(n_img, 8), threads)
    for _ in range(args.steps):
        ips, dt = cpu_port_images_per_s(args.patch, args.prompts, n_img, threads)
        t_all += dt
        n_all += n_img
    val = n_all / t_all
    sample = f"{n_img} images of the same workload per step (oracle port of model.py:219-236 + clip.py:66-79, fp32)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"clip_vitb{args.patch}_zero_shot_ad_224px_{args.prompts}prompts", "batch_per_gpu": n_img},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def side_metrics(dev, pk):
    """HSC head GB/s and AUC ms per 1M scores (the other two figures of BASELINE.json's metric), outside the timed region."""
    from eoe_b200 import metrics, ops
    out = {}
    n, d = 1 << 21, 256
    z = 0.05 * torch.randn(n, d, device=dev)
    y = torch.randint(0, 2, (n,), device=dev)
    for _ in range(3):
        ops.hsc_fused(z, y, 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.hsc_fused(z, y, 0)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    bytes_ = 2 * n * d * 4 + 12 * n + 4
    out["hsc_head"] = {"gbs": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / pk["hbm"], "n": n, "d": d, "dtype": "f32",
                       "note": "fused loss+grad+score, 2.1 GB working set > L2"}
    x = torch.randn(1 << 26, device=dev)
    yb = torch.randint(0, 2, (1 << 26,), device=dev)
    for _ in range(3):
        ops.bce_fused(x, yb, 0)
    a.record()
    for _ in range(10):
        ops.bce_fused(x, yb, 0)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    out["bce_head"] = {"gbs": 20 * (1 << 26) / ms / 1e6, "frac_hbm": 20 * (1 << 26) / ms / 1e6 / pk["hbm"], "n": 1 << 26}
    del x, yb, z, y
    n = 1_000_000
    s = 1 - torch.exp(-torch.randn(n, device=dev).abs())
    yl = (torch.rand(n, device=dev) < 0.5).long()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ws = metrics.AucWorkspace()
    ts = []
    for i in range(8):
        flush.zero_()
        a.record()
        metrics.roc_auc_device(s, yl, workspace=ws)
        b.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    out["auc"] = {"ms_per_1m_scores": statistics.median(ts), "n": n, "note": "device radix sort + scans, L2 flushed"}
    # ---- latency at the sizes the reference itself runs (n = 256 rows per step, 3 000 - 10 000 scores per AUC)
    import numpy as np

    def us(fn, iters=30, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        tt = []
        for _ in range(iters):
            a.record()
            fn()
            b.record(); torch.cuda.synchronize()
            tt.append(a.elapsed_time(b) * 1e3)
        return statistics.median(tt)
    lat = {}
    try:
        from sklearn.metrics import auc as sk_auc, roc_curve as sk_roc
    except Exception:
        sk_auc = sk_roc = None
    for n in (3000, 10000):
        rng = np.random.default_rng(n)
        s_np = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
        y_np = (rng.random(n) < 0.5).astype(np.int64)
        sd_, yd_ = torch.from_numpy(s_np).to(dev), torch.from_numpy(y_np).to(dev)
        ent = {"call_us": us(lambda: metrics.roc_auc_device(sd_, yd_, workspace=ws)), "launches": 1,
               "note": "one call between CUDA events: Python binding + launch + the single kernel (device time alone: "
                       "profiles/r2_latency_reference_sizes.jsonl, device_us_graph_replay)"}
        if sk_auc is not None:
            sk_auc(*sk_roc(y_np, s_np)[:2])                  # warm-up (first call imports / allocates)
            t0 = time.perf_counter()
            for _ in range(5):
                ref = sk_auc(*sk_roc(y_np, s_np)[:2])
            ent["sklearn_host_us"] = (time.perf_counter() - t0) / 5 * 1e6
            ent["bit_exact_vs_sklearn"] = bool(metrics.roc_auc(sd_, yd_) == ref)
        lat[f"auc_n{n}"] = ent
    z = 0.05 * torch.randn(256, 256, device=dev)
    y = (torch.arange(256, device=dev) >= 128).long()
    lat["hsc_fwd_bwd_score_n256_us"] = us(lambda: ops.hsc_fused(z, y, 0))
    x = torch.randn(256, 1, device=dev)
    lat["bce_fwd_bwd_score_n256_us"] = us(lambda: ops.bce_fused(x, y, 0))
    out["latency_at_reference_sizes"] = lat
    return out


def run_ours(args):
    from eoe_b200 import _lib, dist as edist, metrics
    from eoe_b200.encoder import ClipImageEncoder
    from eoe_b200.synth import random_vit_state_dict
    import torch.distributed as tdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the eoe_b200 hot path has no CPU fallback")
    rank, local, ws = edist.init_from_env()
    if ws != args.gpus:
        if ws == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    dev = torch.device("cuda", local)
    numa_cpus = edist.bind_to_gpu_numa(local) if ws > 1 else None      # pinned host buffers land on the GPU's NUMA node
    pk = peaks()
    if args.batch is None:
        args.batch = 512 if args.patch == 16 else 1514
    B, K, S, W, P = args.batch, args.prompts, args.steps, args.warmup, args.patch
    OPS = {"bf16": torch.bfloat16, "f16": torch.float16, "f16x2": "f16x2"}
    op_dtype = OPS[args.dtype]
    mma_mult = 3.0 if args.dtype == "f16x2" else 1.0      # split fp16 pairs: hi*hi + lo*hi + hi*lo per product

    sd = random_vit_state_dict(P, seed=0)
    enc = ClipImageEncoder(sd, device=dev, operand_dtype=op_dtype, max_batch=B, fold_layernorm=not args.no_fold)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    imgs = [torch.randn(B, 3, 224, 224, device=dev, generator=g) for _ in range(2)]     # 2 x 308 MB at B=512: > L2
    text = torch.nn.functional.normalize(torch.randn(K, 512, device=dev, generator=g), dim=-1)
    SW = max(S, W)
    labels = (torch.rand(SW * B, device=dev, generator=g) < 0.9).long()                  # 90 % anomalous
    scores = torch.empty(SW * B, dtype=torch.float32, device=dev)
    auc_ws = metrics.AucWorkspace().ensure(SW * B * ws, dev)

    def job(n_steps, sc, lb, mid=None):
        for k in range(n_steps):
            enc.score(imgs[k & 1], text, out=sc[k * B:(k + 1) * B])
        if mid is not None:
            mid.record()                       # this rank's own scoring is done; what follows waits for the slowest rank
        sc, lb = sc[:n_steps * B], lb[:n_steps * B]
        s_all, l_all = (edist.all_gather_rows(sc, total=ws * n_steps * B), edist.all_gather_rows(lb, total=ws * n_steps * B)) if ws > 1 else (sc, lb)
        return metrics.roc_auc_device(s_all, l_all, workspace=auc_ws)

    def sync():
        if ws > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    job(W, scores, labels)
    sync()
    # Settle: the power-capped clock keeps falling for about a second after load starts (1.6 -> 1.3 GHz), so a timed region
    # that begins right after three warm-up steps runs partly on burst clocks while everything measured after it (the
    # instrumented pass, the e2e passes) does not.  An untimed second of the same job first puts `value`, the roofline
    # pass and `e2e` into the same sustained state -- the state MEASURED_PEAKS.json's sustained figure describes.
    def settle(run):
        """`run(n)` repeated for about args.settle_s seconds; the repeat count is agreed over the ranks (the job contains
        collectives, so every rank must run it the same number of times)."""
        if args.settle_s <= 0:
            return
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        run(min(S, 10))
        b_.record()
        sync()
        tt_ = torch.tensor([a_.elapsed_time(b_)], dtype=torch.float64, device=dev)
        if ws > 1:
            tdist.all_reduce(tt_, op=tdist.ReduceOp.MAX)
        reps = int(args.settle_s * 1e3 / max(float(tt_.item()), 1e-3))
        for _ in range(min(reps, 200)):
            run(min(S, 10))
        sync()

    settle(lambda n_: job(n_, scores, labels))

    # ---- timed region (device-resident inputs).  Pass 1 is clean and gives `value`; pass 2 repeats the identical K
    # steps with every GEMM launch bracketed by a CUDA event pair (98 events per step) and gives the roofline.
    launches0 = _lib.lib().eoe_launch_count()
    sync()
    clocks.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_mid = torch.cuda.Event(enable_timing=True)
    e0.record()
    auc_out, auc_info, _ = job(S, scores, labels, mid=e_mid)
    e1.record()
    sync()
    clocks.end()
    ms = e0.elapsed_time(e1)
    ms_mid = e0.elapsed_time(e_mid)            # this rank's K scoring steps alone (e0 is re-recorded below)
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.lib().eoe_launch_count() - launches0
    auc_val = float(auc_out[0].item())
    enc.profile(True)
    sync()
    e0.record()
    job(S, scores, labels)
    e1.record()
    sync()
    ms_prof = e0.elapsed_time(e1)
    prof = enc.profile_read()
    enc.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank_ms = [ms_mid / S]
    if ws > 1:
        # every rank's own device time for its K scoring steps (before the all-gather, which waits for the slowest rank):
        # `value` is bounded by the MAX of these -- the list shows how far the power-capped GPUs of one box are apart
        allt = torch.zeros(ws, dtype=torch.float64, device=dev)
        tdist.all_gather_into_tensor(allt, torch.tensor([ms_mid], dtype=torch.float64, device=dev))
        per_rank_ms = [float(v) / S for v in allt.cpu().tolist()]
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = ws * S * B / (ms_max / 1e3)

    # ---- e2e: the same job through the public API with HOST (pinned) image batches and host score reads
    host = [torch.randn(B, 3, 224, 224).pin_memory() for _ in range(2)]
    dbuf = [torch.empty(B, 3, 224, 224, device=dev) for _ in range(2)]
    h_scores = torch.empty(S, B, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    e2e_enc = [enc]                            # the encoder the e2e job runs (swapped for the precise-mode side metric)

    def e2e_job(n_steps, host, dbuf):
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        with torch.cuda.stream(copy_stream):
            dbuf[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for k in range(n_steps):
            cur, nxt = k & 1, (k + 1) & 1
            if k + 1 < n_steps:
                with torch.cuda.stream(copy_stream):
                    if k >= 1:
                        copy_stream.wait_event(freed[nxt])
                    dbuf[nxt].copy_(host[nxt], non_blocking=True)       # H2D of step k+1 overlaps compute of step k
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            sc = scores[k * B:(k + 1) * B]
            e2e_enc[0].score(dbuf[cur], text, out=sc)
            freed[cur].record(main)
            h_scores[k % S].copy_(sc, non_blocking=True)                # D2H of the step's result
        s_all, l_all = (edist.all_gather_rows(scores[:n_steps * B], total=ws * n_steps * B),
                        edist.all_gather_rows(labels[:n_steps * B], total=ws * n_steps * B)) \
            if ws > 1 else (scores[:n_steps * B], labels[:n_steps * B])
        out, _, _ = metrics.roc_auc_device(s_all, l_all, workspace=auc_ws)
        return out.cpu()                                                 # the AUC is read on the host (sync)

    e2e_diag = {}

    def time_e2e(hbufs, dbufs, tag=None):
        sampler = None
        if tag == "u8" and rank == 0:          # SM clock / power during the headline e2e region (it runs seconds after `value`)
            sampler = ClockSampler(local)
            sampler.start()
        e2e_job(min(W, S), hbufs, dbufs)
        settle(lambda n_: e2e_job(n_, hbufs, dbufs))    # same settle as before `value`: the host buffers above were built
                                                        # while the GPU idled, so the clocks are back at burst
        sync()                                 # (barrier: every rank enters the timed region together)
        if sampler is not None:
            sampler.begin()
        t0 = time.perf_counter()
        e0.record()
        e2e_job(S, hbufs, dbufs)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if sampler is not None:
            sampler.end()
            e2e_diag["u8_clocks"] = sampler.stop()
        tt = torch.tensor([wall, e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
        if ws > 1:
            tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        if tag:       # wall clock (what `value` of the e2e block is computed from) next to the device time of the same region
            e2e_diag[tag] = {"wall_ms_per_step": float(tt[0]) * 1e3 / S, "device_ms_per_step": float(tt[1]) * 1e3 / S}
        return ws * S * B / float(tt[0])

    # ---- headline e2e: the dataset's decoded pixels as they sit in host memory -- uint8 NHWC at 224 x 224 -- with ToTensor
    # + Normalize fused into the patchify kernel (eoe_vit_encode_u8)
    host8 = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dbuf8 = [torch.empty(B, 224, 224, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    e2e_u8_val = time_e2e(host8, dbuf8, "u8")
    del host8, dbuf8
    # ---- the same job from fp32 NCHW host batches (what the reference's DataLoader hands over after ToTensor + Normalize)
    e2e_val = time_e2e(host, dbuf, "f32")
    del host, dbuf

    # ---- e2e from RAW decoded images of the dataset's native size: Resize(bicubic) + CenterCrop + ToTensor + Normalize
    # (all of CLIP's `_transform`) run inside the patchify kernel (eoe_vit_encode_u8_resize)
    rh, rw = (32, 32) if P == 32 else (375, 500)              # CIFAR-10-shaped (config 2) / ImageNet-shaped (config 3)
    hostr = [torch.randint(0, 256, (B, rh, rw, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dbufr = [torch.empty(B, rh, rw, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    e2e_raw_val = time_e2e(hostr, dbufr)
    del hostr, dbufr

    # ---- the other operand dtype, device-resident, same K steps (side metric; parity of both: DESIGN.md section 5)
    other, others = None, {}
    for o_name in ([] if args.no_side else [n for n in ("bf16", "f16", "f16x2") if n != args.dtype]):
        if o_name == "f16x2" and args.no_fold:
            continue
        enc_o = ClipImageEncoder(sd, device=dev, operand_dtype=OPS[o_name], max_batch=B, fold_layernorm=not args.no_fold)

        def job_o(n_steps):
            for k in range(n_steps):
                enc_o.score(imgs[k & 1], text, out=scores[k * B:(k + 1) * B])
        job_o(W)
        sync()
        settle(job_o)
        e0.record()
        job_o(S)
        e1.record()
        sync()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if ws > 1:
            tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        others[o_name] = {"dtype": o_name, "value": ws * S * B / (float(tt.item()) / 1e3), "unit": UNIT,
                          "note": "same K steps, device-resident, encoder + fused score only (no all-gather / AUC)"}
        if o_name == "f16x2":
            enc_o.profile(True)
            job_o(S)
            sync()
            pr = enc_o.profile_read()
            enc_o.profile(False)
            g_ms_o, g_fl_o = sum(v[0] for v in pr.values()), sum(v[2] for v in pr.values())
            others[o_name].update({
                "what": "PRECISE mode: every stored 16-bit tensor is an fp16 (hi, lo) pair, products = hi*hi + lo*hi + hi*lo "
                        "(3x the MMA work); end-to-end scores within 1e-3 relative of the fp32 reference on EVERY image "
                        "(tests/test_gpu_encoder_split.py::test_end_to_end_scores_split: max 6.6e-5 / 3.4e-5)",
                "gemm_mma_tflops_executed": 3.0 * g_fl_o / (g_ms_o * 1e-3) / 1e12 if g_ms_o > 0 else None,
                "gemm_frac_of_sustained_peak": 3.0 * g_fl_o / (g_ms_o * 1e-3) / 1e12 / pk["tf_sust"] if g_ms_o > 0 else None,
                "gemm_share_of_step": g_ms_o / max(float(tt.item()), 1e-9)})
            # the same end-to-end job as the headline `e2e` (pinned uint8 host batches in, scores + AUC out) in precise mode
            host8 = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
            dbuf8 = [torch.empty(B, 224, 224, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
            e2e_enc[0] = enc_o
            others[o_name]["e2e"] = {"value": time_e2e(host8, dbuf8), "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224,
                                     "d2h_bytes_per_step": B * 4 + 8}
            e2e_enc[0] = enc
            del host8, dbuf8
        del enc_o
        torch.cuda.empty_cache()
    other = others.get("bf16" if args.dtype == "f16" else "f16")

    # ---- data-parallel TRAINING at model scale (BASELINE configs 4 / 5): step, all-reduce and overlap
    # of a ResNet-18-sized BCE run and a ViT-B/16-sized HSC run through the fused head kernels + eoe_b200.dist.GradBuckets
    dp_rows = None
    if not args.no_side:
        del imgs, scores
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import dp_bench
        _, _, dp_rows = dp_bench.run(("cfg4", "cfg5"), steps=5, warmup=2)

    if rank != 0:
        if ws > 1:
            tdist.barrier()
            tdist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the tcgen05 GEMM: all launches of the timed region)
    g_ms = sum(v[0] for v in prof.values())
    g_fl = sum(v[2] for v in prof.values())
    g_n = sum(v[1] for v in prof.values())
    achieved = mma_mult * g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0     # EXECUTED tensor flops (f16x2: 3 MMAs per product)
    traffic, traffic_note = None, None
    build_id = _lib.lib().eoe_build_id().decode()
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        ent = (tj.get("by_dtype") or {}).get(args.dtype, tj)
        from eoe_b200 import build as ebuild
        # the capture is valid for this library if it IS the profiled build, or if this library was built from the tree's
        # sources and the sources the GEMM kernels come from (gemm_sm100.cuh, sm100_ptx.cuh, common.cuh, vit.cu, flags) are
        # byte-identical to the profiled build's (a later change to the heads / AUC translation units does not touch them)
        same_gemm = (tj.get("gemm_source_id") is not None and ebuild.source_id() == build_id
                     and tj.get("gemm_source_id") == ebuild.gemm_source_id())
        if tj.get("build_id") != build_id and not same_gemm:
            # an ncu capture of ANOTHER build says nothing about this one: refuse it instead of quoting a stale figure
            traffic_note = {"refused": f"profiles/gemm_traffic.json was captured on build {tj.get('build_id')}, "
                                       f"this library is build {build_id}"}
        elif tj.get("batch") == B and P == 16 and ent.get("dram_bytes_per_launch"):
            traffic = ent.get("dram_bytes_per_launch")         # ncu --set full capture of the c_fc GEMM at this batch size
            traffic_note = {"kernel": ent.get("kernel"), "algorithmic_bytes_per_launch": ent.get("algorithmic_bytes_per_launch"),
                            "source": ent.get("source"), "build_id": build_id, "captured_on_build": tj.get("build_id"),
                            "gemm_source_id": tj.get("gemm_source_id"), "note": ent.get("note")}
    fc = prof.get("c_fc", (0.0, 0, 0.0))
    Lt = (224 // P) ** 2 + 1
    exec_gflop = (g_fl / (S * B) + (enc.n_layers - 1) * 4.0 * Lt * Lt * enc.width + 4.0 * Lt * enc.width) / 1e9
    roofline = {
        "bound": "tensor", "kernel": "eoe::gemm::gemm_kernel (tcgen05, all 49 GEMM launches per step)",
        "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
        "traffic": traffic, "traffic_of": traffic_note,
        "dominant_instance": {"kernel": "gemm_kernel<ln_2 fold + bias + 1.702*QuickGELU> (c_fc)", "launches": fc[1],
                              "achieved": (fc[2] / (fc[0] * 1e-3) / 1e12 if fc[0] > 0 else None), "unit": "TFLOP/s",
                              "flops_per_launch": (fc[2] / fc[1] if fc[1] else None), "us_per_launch": (1e3 * fc[0] / fc[1] if fc[1] else None)},
        "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
        "launches": g_n, "gemm_ms_per_step": g_ms / S, "gemm_share_of_step": g_ms / ms_prof,
        "ms_per_step_instrumented": ms_prof / S,
        "per_kind_tflops": {k: (v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else None) for k, v in prof.items()},
        "encoder_tensor_frac": value / ws * GFLOP_PER_IMG[P] * 1e9 / 1e12 / pk["tf_sust"],
        "encoder_tensor_frac_burst": value / ws * GFLOP_PER_IMG[P] * 1e9 / 1e12 / pk["tf_burst"],
        # the last block is evaluated for the class-token rows only (identical output): EXECUTED flops per image =
        # GEMM launches of the timed region + attention (full L x L in all but the last block, one query row there)
        "executed_gflop_per_image": exec_gflop, "algorithmic_gflop_per_image": GFLOP_PER_IMG[P],
        "encoder_tensor_frac_executed": value / ws * exec_gflop * 1e9 / 1e12 / pk["tf_sust"],
        "encoder_tensor_frac_executed_burst": value / ws * exec_gflop * 1e9 / 1e12 / pk["tf_burst"],
        "build_id": build_id,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": S, "warmup": W,
        "ms_per_step": ms_max / S, "per_rank_scoring_ms_per_step": per_rank_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"clip_vitb{P}_zero_shot_ad_224px_{K}prompts", "batch_per_gpu": B, "global_batch": B * ws,
                   "images_scored": ws * S * B, "weights": "random-init ViT-B/%d visual tower (reference state_dict layout)" % P,
                   "l2": "inputs exceed L2 (2 alternating %.0f MB image batches per GPU)" % (B * 3 * 224 * 224 * 4 / 1e6),
                   "parallelism": f"dp{ws}: images sharded by rank, all_gather(scores, labels) -> global device AUC",
                   "timed_region": "K x (patchify + ViT encoder + fused score head) + all-gather + ROC-AUC",
                   "settle": f"{args.settle_s} s of the same job, untimed, after the W warm-up steps (sustained power-capped clocks)"},
        "roofline": roofline,
        "e2e": {"value": e2e_u8_val, "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224, "d2h_bytes_per_step": B * 4 + 8,
                "note": "public API (ClipImageEncoder.score + metrics.roc_auc_device) from pinned HOST batches of decoded uint8 "
                        "NHWC pixels (ToTensor + Normalize fused into the patchify kernel, eoe_vit_encode_u8), double-buffered "
                        "H2D on a copy stream, every step's scores and the final AUC read back to the host",
                "numa_cpus": numa_cpus, "timing": e2e_diag.get("u8"), "clocks": e2e_diag.get("u8_clocks")},
        "e2e_f32": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224 * 4, "d2h_bytes_per_step": B * 4 + 8,
                    "note": "same job from pinned fp32 NCHW host batches (the reference DataLoader's output format): 4x the PCIe bytes",
                    "timing": e2e_diag.get("f32")},
        "e2e_raw": {"value": e2e_raw_val, "unit": UNIT, "h2d_bytes_per_step": B * rh * rw * 3, "d2h_bytes_per_step": B * 4,
                    "note": f"same job from raw uint8 {rh}x{rw} images: Resize(bicubic, Pillow-exact) + CenterCrop + ToTensor + "
                            "Normalize fused into the patchify kernel (eoe_vit_encode_u8_resize)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "auc": auc_val,
    }
    if ws == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_img = args.cpu_baseline_images
        ips, dt = cpu_port_images_per_s(P, K, n_img, threads)
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n_img} images of the same workload, {dt:.1f} s (oracle port, fp32, torch CPU)"}
    if not args.no_side:
        line["side_metrics"] = side_metrics(dev, pk)
        line["side_metrics"]["other_dtype"] = other
        line["side_metrics"]["precise_f16x2"] = others.get("f16x2")
        line["side_metrics"]["other_dtypes"] = others
        line["side_metrics"]["dp_train"] = dp_rows
    emit(line)
    if ws > 1:
        tdist.barrier()
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None,
                    help="images per GPU per step; default 512 (ViT-B/16: 394 m-tiles of 256 rows = 16 waves of 74 CTA pairs at "
                         "N = 768) or 1514 (ViT-B/32: 296 m-tiles = 4 x 74; 512 images leave the N = 768 GEMMs at 4.05 waves)")
    ap.add_argument("--patch", type=int, default=16, choices=[16, 32])
    ap.add_argument("--prompts", type=int, default=30)
    ap.add_argument("--dtype", default="f16", choices=["bf16", "f16", "f16x2"],
                    help="GEMM operand dtype.  f16 (default) is the reference's own GPU dtype and the one whose end-to-end scores "
                         "are closer to the fp32 reference than the reference's GPU path (DESIGN.md section 5); bf16 is ~5 %% faster")
    ap.add_argument("--ref-images", type=int, default=64, help="images per step of the CPU port (bounded sample)")
    ap.add_argument("--cpu-baseline-images", type=int, default=384,
                    help="bounded sample of the cpu_baseline leg inside our arm (about 10 s of CPU work on 16 cores)")
    ap.add_argument("--settle-s", type=float, default=1.0,
                    help="untimed seconds of the same job between the warm-up steps and the timed region (power-capped steady state)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true")
    ap.add_argument("--no-fold", action="store_true", help="stand-alone ln_1 / ln_2 kernels instead of the LayerNorm fold (A/B)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3            # timing rule: at least 3 warm-up steps
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner to fd 1
    # when NCCL_DEBUG=VERSION is set on the box), so fd 1 points at stderr until the line is printed.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
