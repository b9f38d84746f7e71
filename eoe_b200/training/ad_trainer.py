"""Re-hosted trainer runtime for the hot path (reference: src/eoe/training/ad_trainer.py).

Kept from the reference: the three abstract hooks (ad_trainer.py:624-662) with identical names / arguments /
kwargs (`inputs=imgs, nominal_label=...`, :430,434-436,508), the batch loop of `train_cls` (:406-444), the
NaN guard (:447-449, `NanGradientsError`), the per-epoch training AUC (:452-455), `eval_cls` (:498-527) and
its `(ROC, PRC)` result, optimiser / scheduler choice (:379-384).

Changed for B200: scores and labels stay on the device (the reference does `.cpu()` per batch, :436-440,
509-511); one device AUC per epoch (eoe_b200.metrics, bit-exact with sklearn) instead of sklearn on the host;
one host sync per epoch instead of three per batch; optional data-parallel training / sharded evaluation over
torch.distributed (new; see eoe_b200.dist).  Out of scope: datasets, transforms, logging,
snapshots -- `train_cls` / `eval_cls` take a loader yielding the reference's batch triple `(imgs, lbls, idcs)`
(datasets/bases.py:591-597) instead of a dataset object.
"""
import weakref
from abc import ABC, abstractmethod
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from .. import dist as edist
from .. import metrics


class NanGradientsError(RuntimeError):
    """ad_trainer.py:23-24: raised when an epoch produced NaN anomaly scores."""


class ADTrainer(ABC):
    AD_MODES = ("one_vs_rest", "leave_one_out")

    def __init__(self, model: Optional[torch.nn.Module], epochs: int = 0, lr: float = 1e-3, wdk: float = 0.0,
                 milestones: Sequence[int] = (), batch_size: int = 128, ad_mode: str = "one_vs_rest",
                 device="cuda", data_parallel: bool = False, sgd: bool = False, graph_step: bool = False):
        if ad_mode not in self.AD_MODES:
            raise NotImplementedError(f"AD mode {ad_mode} unknown. Known modes are {self.AD_MODES}.")
        self.model = model
        self.epochs, self.lr, self.wdk, self.milestones = epochs, lr, wdk, list(milestones)
        self.batch_size, self.ad_mode = batch_size, ad_mode
        self.device = torch.device(device)
        self.center = None
        self.data_parallel = data_parallel
        self.sgd = sgd                       # the reference uses SGD(nesterov) iff the model is CLIP (:379-383)
        # graph_step: capture one whole training step (model forward, fused loss + score kernel, backward, optimiser)
        # into a CUDA graph and replay it per batch: at the reference's batch size (128 + 128 rows of a
        # small CNN) the step is ~100 kernel launches of a few microseconds each, i.e. bound by the host.
        self.graph_step = graph_step
        self._score_cache = None

    # ---------------------------------------------------------------- hooks (ad_trainer.py:624-662)
    @abstractmethod
    def prepare_metric(self, cstr: str, loader, model: torch.nn.Module, seed: int, **kwargs) -> torch.Tensor:
        pass

    @abstractmethod
    def compute_anomaly_score(self, features: torch.Tensor, center: torch.Tensor, **kwargs) -> torch.Tensor:
        pass

    @abstractmethod
    def loss(self, features: torch.Tensor, labels: torch.Tensor, center: torch.Tensor, **kwargs) -> torch.Tensor:
        pass

    # scores computed by the fused loss kernel are handed to the compute_anomaly_score call that follows on
    # the same feature tensor (ad_trainer.py:430-436 calls loss, then compute_anomaly_score, on `image_features`)
    # The entry holds a weak reference to that very tensor object (an address could be recycled by the caching allocator
    # for the next same-shape batch) plus its version counter, is consumed by the first lookup and dropped by any miss.
    def _remember_scores(self, features, scores, tag=None):
        self._score_cache = (weakref.ref(features), features._version, tag, scores)

    def _cached_scores(self, features, tag=None):
        c, self._score_cache = self._score_cache, None
        if c is not None and c[0]() is features and c[1] == features._version and c[2] == tag:
            return c[3]
        return None

    def _sync_metric(self, center):
        """Data-parallel runs: every rank must optimise the SAME objective, so the tensor prepare_metric returned on
        rank 0 (computed from rank 0's shard: e.g. the DSVDD centre, dsvdd.py:11-21) is broadcast to all ranks."""
        _, ws = edist.world()
        if ws > 1 and self.data_parallel and torch.is_tensor(center):
            center = center.contiguous()
            torch.distributed.broadcast(center, src=0)
        return center

    def _check_equal_batches(self, loader):
        """Gradient buckets average over ranks (equal weights) and every rank must join every all-reduce: per-rank loaders
        need the same number of batches (and equal batch sizes for the average to be the global-batch mean)."""
        _, ws = edist.world()
        if ws > 1 and self.data_parallel and hasattr(loader, "__len__"):
            n = torch.tensor([len(loader)], dtype=torch.int64, device=self.device)
            lo, hi = n.clone(), n.clone()
            torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
            torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
            if int(lo) != int(hi):
                raise ValueError(f"data-parallel training needs the same number of batches on every rank (got {int(lo)}..{int(hi)})")

    # ---------------------------------------------------------------- training (ad_trainer.py:356-471)
    def train_cls(self, model: torch.nn.Module, loader: Iterable, nominal_label: int = 0, clsstr: str = "",
                  seed: int = 0, epochs: Optional[int] = None) -> Tuple[torch.nn.Module, Optional[metrics.ROC], List[float]]:
        model = model.to(self.device).train()
        self._score_cache = None
        epochs = self.epochs if epochs is None else epochs
        params = [p for p in model.parameters() if p.requires_grad]
        _, ws = edist.world()
        if epochs == 0 or not params:
            # zero-shot (ad_trainer.py:406 never runs for epochs == 0): only prepare_metric is executed.  The B200
            # image encoder is inference-only, so it has no trainable parameters to hand to an optimiser.
            if epochs > 0:
                raise ValueError("train_cls with epochs > 0 needs a model with trainable parameters")
            self.center = self.prepare_metric(clsstr, loader, model, seed)
            return model.eval(), None, []
        use_graph = self.graph_step and self.device.type == "cuda" and not (self.data_parallel and ws > 1)
        # a captured step reads the learning rate from a device tensor, which MultiStepLR updates in place (Adam:
        # capturable=True; SGD: the fused implementation takes a tensor lr -- the foreach one would call .item() on it)
        lr = torch.tensor(self.lr, dtype=torch.float32, device=self.device) if use_graph else self.lr
        if self.sgd:
            opt = torch.optim.SGD(params, lr=lr, weight_decay=self.wdk, momentum=0.9, nesterov=True,
                                  **({"fused": True} if use_graph else {}))
        else:
            opt = torch.optim.Adam(params, lr=lr, weight_decay=self.wdk, amsgrad=False, capturable=use_graph)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, self.milestones, 0.1)
        buckets = edist.GradBuckets(params) if (self.data_parallel and ws > 1) else None
        center = self.center = self._sync_metric(self.prepare_metric(clsstr, loader, model, seed))
        self._check_equal_batches(loader)
        cls_roc, losses = None, []
        graph = None            # (CUDAGraph, static imgs, static lbls, static loss, static scores, lr) once captured
        for ep in range(epochs):
            ep_labels, ep_ascores, ep_losses = [], [], []
            for imgs, lbls, _idcs in loader:
                imgs = imgs.to(self.device, non_blocking=True)
                lbls = lbls.to(self.device, non_blocking=True)
                if use_graph:
                    lr_now = opt.param_groups[0]["lr"]
                    if graph is not None and not (torch.is_tensor(lr_now) and lr_now is graph[5]):
                        # the schedule replaced the lr object the graph reads (or it is a python float baked into it)
                        if not (isinstance(lr_now, float) and isinstance(graph[5], float) and lr_now == graph[5]):
                            graph = None                                   # capture again with the current lr
                    if graph is None:
                        graph = self._capture_step(model, opt, center, imgs, lbls, nominal_label)
                    if graph[1].shape == imgs.shape and graph[2].shape == lbls.shape:
                        graph[1].copy_(imgs)
                        graph[2].copy_(lbls)
                        graph[0].replay()
                        ep_labels.append(lbls.detach())
                        ep_ascores.append(graph[4].clone())
                        ep_losses.append(graph[3].clone())
                        continue
                    # a ragged last batch runs eagerly below
                if buckets is not None:
                    buckets.zero_grad()
                else:
                    opt.zero_grad()
                image_features = model(imgs)
                loss = self.loss(image_features, lbls, center, inputs=imgs, nominal_label=nominal_label)
                loss.backward()
                if buckets is not None:
                    buckets.finish()
                opt.step()
                anomaly_scores = self.compute_anomaly_score(image_features, center, inputs=imgs, nominal_label=nominal_label)
                ep_labels.append(lbls.detach())
                ep_ascores.append(anomaly_scores.detach().reshape(-1))
                ep_losses.append(loss.detach())
            ep_labels, ep_ascores = torch.cat(ep_labels), torch.cat(ep_ascores)
            if ws > 1 and self.data_parallel:
                ep_labels, ep_ascores = edist.all_gather_rows(ep_labels), edist.all_gather_rows(ep_ascores)
            # one host sync per epoch: NaN flag, class presence and the mean loss travel together
            flags = torch.stack([ep_ascores.isnan().any().float(), (ep_labels == 1).any().float(),
                                 torch.stack(ep_losses).float().mean()]).cpu()
            if flags[0] > 0:
                raise NanGradientsError()                                 # ad_trainer.py:448-449
            losses.append(float(flags[2]))
            if flags[1] > 0:                                              # ad_trainer.py:452-455
                roc, _ = metrics.roc_curve_auc(ep_ascores, ep_labels)
                cls_roc = roc if roc is not None else cls_roc
            sched.step()
        return model.eval(), cls_roc, losses

    def _capture_step(self, model, opt, center, imgs, lbls, nominal_label):
        """Capture `model -> loss (fused kernel) -> backward -> optimiser step -> scores` once, on static input buffers.
        Three eager steps run on a side stream first (PyTorch's whole-network capture recipe: lazy initialisations and the
        optimiser's state tensors must exist before capture).  They must not count as training: weights, BatchNorm
        statistics and optimiser state are snapshotted before and restored in place afterwards."""
        import copy
        s_imgs, s_lbls = imgs.clone(), lbls.clone()
        saved_model = copy.deepcopy(model.state_dict())
        saved_opt = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in opt.state.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                opt.zero_grad(set_to_none=True)
                f = model(s_imgs)
                l = self.loss(f, s_lbls, center, inputs=s_imgs, nominal_label=nominal_label)
                l.backward()
                opt.step()
                self.compute_anomaly_score(f, center, inputs=s_imgs, nominal_label=nominal_label)
        torch.cuda.current_stream(self.device).wait_stream(side)
        # undo the warm-up IN PLACE (the optimiser's state tensors must exist before capture, or they would be created --
        # and re-zeroed at every replay -- inside the graph): weights, BatchNorm statistics, moments and step counters
        # return to their values before it (zero where the state did not exist yet)
        model.load_state_dict(saved_model)
        for p, st in opt.state.items():
            for k, v in st.items():
                if torch.is_tensor(v):
                    old = saved_opt.get(p, {}).get(k)
                    if torch.is_tensor(old):
                        v.copy_(old)
                    else:
                        v.zero_()
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            f = model(s_imgs)
            s_loss = self.loss(f, s_lbls, center, inputs=s_imgs, nominal_label=nominal_label)
            s_loss.backward()
            opt.step()
            s_scores = self.compute_anomaly_score(f, center, inputs=s_imgs, nominal_label=nominal_label).detach().reshape(-1)
        return g, s_imgs, s_lbls, s_loss.detach(), s_scores, opt.param_groups[0]["lr"]

    # ---------------------------------------------------------------- evaluation (ad_trainer.py:473-550)
    @torch.no_grad()
    def eval_cls(self, model: torch.nn.Module, loader: Iterable, nominal_label: int = 0, clsstr: str = "",
                 gather: bool = True) -> Tuple[Optional[metrics.ROC], Optional[metrics.PRC]]:
        """`loader` yields this rank's shard of the test set; with torch.distributed initialised and gather=True the
        per-rank score / label shards are all-gathered (rank order == index order for eoe_b200.dist.shard_range)
        and every rank computes the global ROC / PRC."""
        model = model.to(self.device).eval() if isinstance(model, torch.nn.Module) else model
        self._score_cache = None
        center = self.center
        ep_labels, ep_ascores = [], []
        for imgs, lbls, _idcs in loader:
            imgs = imgs.to(self.device, non_blocking=True)
            image_features = model(imgs)
            anomaly_scores = self.compute_anomaly_score(image_features, center, inputs=imgs, nominal_label=nominal_label)
            ep_labels.append(lbls.to(self.device, non_blocking=True))
            ep_ascores.append(anomaly_scores.reshape(-1))
        ep_labels, ep_ascores = torch.cat(ep_labels), torch.cat(ep_ascores)
        if gather:
            ep_labels, ep_ascores = edist.all_gather_rows(ep_labels), edist.all_gather_rows(ep_ascores)
        self.last_eval = (ep_labels, ep_ascores)
        # ad_trainer.py:516-527: both classes must be present, labels < 0 are dropped, else (None, None)
        return metrics.roc_curve_auc(ep_ascores, ep_labels, with_prc=True, ignore_negative_labels=True)
