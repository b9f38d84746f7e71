"""Multi-GPU plumbing (new functionality: the reference is single-GPU, src/eoe/main/__init__.py:110-114).

One process per GPU, `torch.distributed` over NCCL/NVLink (gloo on CPU for the host-logic tests).  The hot path
shards by rows (images / score rows are independent), so there are exactly two collectives:
  * all_gather of per-rank score / label shards before the global AUC      -> `all_gather_rows`
  * bucketed all-reduce of gradients, overlapped with backward               -> `GradBuckets`
"""
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bind_to_gpu_numa(local_rank: int) -> Optional[List[int]]:
    """Pin this process to the CPUs local to its GPU's PCIe root (sysfs `local_cpulist`), so that pinned host buffers
    allocated afterwards are first-touched on the GPU's NUMA node and the H2D feed does not cross the socket interconnect
    (the e2e feed of N ranks on one host).  Returns the CPU list, or None when the topology is not
    exposed (single-node VMs): then nothing is changed."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        cpus: List[int] = []
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT. Returns (rank, local_rank, world)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if be == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=be, rank=rank, world_size=ws, **kw)
    return rank, local, ws


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of n rows: rank r owns rows r*ceil(n/W) ... (last shards may be short/empty)."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_rows(t: torch.Tensor, group=None, total: Optional[int] = None) -> torch.Tensor:
    """Concatenate 1-D (or [n, ...]) per-rank shards of possibly different length, in rank order, on every rank.
    `total`: the global row count when the shards follow `shard_range(total, rank, world)` -- the shard sizes are then
    known on every rank and the size exchange (a collective plus a host sync) is skipped."""
    rank, ws = world()
    if ws == 1:
        return t
    t = t.contiguous()
    if total is not None:
        sizes = [hi - lo for lo, hi in (shard_range(total, r, ws) for r in range(ws))]
        if sizes[rank] != t.shape[0]:
            raise ValueError(f"rank {rank} holds {t.shape[0]} rows, shard_range({total}, {rank}, {ws}) says {sizes[rank]}")
    else:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        gathered = torch.zeros(ws, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(gathered, n, group=group)
        sizes = [int(v) for v in gathered.cpu().tolist()]                  # one host sync instead of one per rank
    mx = max(sizes)
    if mx == 0:
        return t
    if t.shape[0] == mx:
        pad = t
    else:
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
    out = torch.empty((ws * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(ws)])


class GradBuckets:
    """Data-parallel gradient averaging: parameters' .grad are views into a few flat buckets (filled in reverse
    parameter order, the order backward produces them); a bucket's all-reduce is launched asynchronously from a
    post-accumulate-grad hook as soon as its last gradient has been written, so communication overlaps the rest of
    backward.  `finish()` waits and scales by 1/world.  Use `zero_grad()` of this object (keeps the views)."""

    def __init__(self, params, bucket_bytes: int = 32 << 20, group=None, tail_bytes: int = 4 << 20):
        self.group = group
        self.enabled = True          # False: gradients stay local (no collective is launched; A/B timing of the overlap)
        self.rank, self.ws = world()
        # NCCL averages inside the collective (ReduceOp.AVG): no extra pass over the buckets after the wait; gloo (CPU
        # tests) only sums, so the division stays a separate step there
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._counts: List[int] = []
        self._works = []
        cur, cur_bytes = [], 0
        groups = []
        for p in reversed(self.params):
            nb = p.numel() * p.element_size()
            if cur and (cur_bytes + nb > bucket_bytes or cur[0].dtype != p.dtype):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            groups.append(cur)
        # The LAST bucket holds the first layers' gradients and can only be launched when backward ends: nothing is left
        # to overlap it with, so it is kept small (its all-reduce is then latency, not bandwidth: tools/dp_bench.py).
        if groups and len(groups[-1]) > 1 and tail_bytes > 0:
            last, tail, nb = groups[-1], [], 0
            while len(last) > 1 and nb + last[-1].numel() * last[-1].element_size() <= tail_bytes:
                nb += last[-1].numel() * last[-1].element_size()
                tail.insert(0, last.pop())
            if tail:
                groups.append(tail)
        for bi, ps in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
            off = 0
            for p in ps:
                p.grad = flat[off: off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[p] = bi
                if self.ws > 1:
                    p.register_post_accumulate_grad_hook(self._hook)
            self.buckets.append(flat)
            self._counts.append(len(ps))
        self._pending = list(self._counts)

    def _hook(self, p):
        if not self.enabled:
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._works.append(dist.all_reduce(self.buckets[bi], op=self._op, group=self.group, async_op=True))

    def finish(self):
        """Call after loss.backward(): wait for the in-flight all-reduces, average, re-arm."""
        if self.ws > 1 and self.enabled:
            for bi, left in enumerate(self._pending):          # parameters that received no gradient this step
                if left != 0:
                    self._works.append(dist.all_reduce(self.buckets[bi], op=self._op, group=self.group, async_op=True))
            for w in self._works:
                w.wait()
            if not self._avg:
                for b in self.buckets:
                    b.div_(self.ws)
        self._works = []
        self._pending = list(self._counts)

    def zero_grad(self):
        for b in self.buckets:
            b.zero_()
