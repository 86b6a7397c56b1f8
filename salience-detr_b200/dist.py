"""Multi-GPU plumbing of the path: batch-sharded independent replicas.

The encoder forward has NO collective: each rank owns its local batch (the reference under DDP behaves the same:
per-process local batch, train_config.py:9).  torch.distributed is used only for (a) the barrier + max-over-ranks
timing of bench.py and (b) the gradient all-reduce of the training configuration."""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None) -> bool:
    """Initialise the default process group from torchrun's environment (MASTER_ADDR defaults to 127.0.0.1)."""
    rank, world, _ = env_rank_world()
    if world <= 1:
        return False
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    # NCCL's version banner / NCCL_DEBUG output goes to stdout by default; callers that print machine-readable results there
    # (bench.py: ONE JSON line) need it on stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return True


def max_over_ranks(value: float, device="cpu") -> float:
    """The slowest rank's time: what a multi-GPU throughput number must be computed from."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch_seed(base_seed: int, rank: int) -> int:
    """Every replica draws its own synthetic batch (weak scaling: per-GPU work is fixed)."""
    return base_seed + rank


def aggregate_throughput(units_per_rank_step: int, steps: int, world: int, total_ms_max: float) -> float:
    """Whole-job units/s = all ranks' units / slowest rank's time."""
    return units_per_rank_step * steps * world / (total_ms_max / 1000.0)


def allreduce_gradients_(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20):
    """Mean gradient all-reduce in flat buckets (the one collective of the training configuration,
    accelerate/DDP in the reference: main.py:144, util/engine.py:58).  Over NVSwitch the cost is per-launch
    latency, not per-link bandwidth, so buckets are sized for few launches."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


class GradientBuckets:
    """Bucketed gradient all-reduce OVERLAPPED with the backward pass (the one collective of the training configuration:
    DDP / accelerate in the reference, main.py:144, util/engine.py:58).

    Parameters are packed, in reverse registration order (~ the order their gradients become ready), into flat fp32
    buckets of about ``bucket_bytes``; every ``p.grad`` is a VIEW into its bucket, so autograd accumulates straight into
    the communication buffer (no flatten / copy-back passes).  A post-accumulate hook counts ready gradients and launches
    the bucket's asynchronous ``all_reduce`` (NCCL over NVLink/NVSwitch on its own stream) as soon as the last one lands,
    while autograd keeps running the backward of earlier layers.  ``finish()`` waits for the handles and turns sums into
    means.  Over NVSwitch the cost is per-launch latency, not per-link bandwidth: few large buckets.

    Usage per step:  buckets.zero_();  loss.backward();  buckets.finish();  optimizer.step()"""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, process_group=None):
        self.group = process_group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.world = dist.get_world_size(process_group) if self.active else 1
        params = [p for p in params if p.requires_grad]
        self.buckets = []      # [(flat buffer, [params])]
        self._bucket_of = {}
        cur, size = [], 0
        for p in reversed(params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._close(cur)
                cur, size = [], 0
        if cur:
            self._close(cur)
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self.launch_order = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]

    def _close(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            self._bucket_of[p] = len(self.buckets)
        self.buckets.append((flat, list(ps)))

    def zero_(self):
        """Start of a step: clear the communication buffers (= all gradients) and the ready counters."""
        for flat, _ in self.buckets:
            flat.zero_()
        self._pending = [len(ps) for _, ps in self.buckets]
        self._handles, self.launch_order = [], []

    def _on_grad(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        self.launch_order.append(i)
        if self.active:
            self._handles.append(dist.all_reduce(self.buckets[i][0], group=self.group, async_op=True))

    def finish(self):
        """After backward: flush buckets whose parameters got no gradient this step, wait, sum -> mean."""
        for i, n in enumerate(self._pending):
            if n > 0:  # some parameter of the bucket was not reached by the loss: its slice is zero on every rank
                self._pending[i] = 0
                self._launch(i)
        for h in self._handles:
            h.wait()
        if self.world > 1:
            for flat, _ in self.buckets:
                flat.div_(self.world)

    @property
    def total_bytes(self) -> int:
        return sum(f.numel() * 4 for f, _ in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and hence its later pinned-host allocations, first touch) to the CPUs of the NUMA node the GPU
    hangs off.  One rank per GPU launched by torchrun otherwise floats over both sockets: on the 8-GPU hosts GPUs 4-7 sit
    on node 1, and round 1 measured the end-to-end (host-buffer) rate falling to 0.73 of the device rate at N = 8.
    Returns what was done (for the bench line); a no-op when sysfs does not say."""
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info.update(pci=bdf, node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:  # sysfs layout / permissions differ: keep running unbound
        info["error"] = f"{type(e).__name__}: {e}"
    return info
