"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

The hot path shards without any data-path collective:

* training      batch elements are independent through encoder and walk (reference model.py:16-46 has no
                cross-batch op but the final mean) -> each rank takes ``B`` radargram items; the only exchange
                is the all-reduce of the ENCODER gradients (the reference's nn.DataParallel does the same
                reduction implicitly, scripts/train.py:45-47).  The walk itself has no parameters.
* propagation   radargrams are independent (each is re-seeded from its own reference column,
                scripts/test/test_all.py:91-95) -> contiguous ranges of radargrams per rank, no collective.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``n_items`` for ``rank``; sizes differ by at most one, earlier ranks larger."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int, group=None, local_items: int = 1,
                        total_items: int = 0) -> None:
    """Combine the gradients of ``params`` across ranks with ONE flat all-reduce (19.9 MB for the ResNet encoder).

    Every parameter that requires a gradient takes part on every rank, in parameter order (a parameter without a gradient on
    this rank contributes zeros), so the flat buffers always line up.  Each rank's gradient is the gradient of ITS mean loss
    over ``local_items`` batch elements; it is weighted by ``local_items / total_items`` so that uneven shards
    (``shard_range`` with B % world != 0) still give the gradient of the full-batch mean loss.  With ``total_items = 0`` the
    shards are taken to be equal (plain average).  Equivalent to DistributedDataParallel with a single bucket.
    """
    if world == 1:
        return
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return
    weight = (float(local_items) / float(total_items)) if total_items else 1.0 / world
    grads = [(p.grad if p.grad is not None else torch.zeros_like(p)) for p in ps]
    flat = torch._utils._flatten_dense_tensors(grads).mul_(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    for p, f in zip(ps, torch._utils._unflatten_dense_tensors(flat, grads)):
        if p.grad is None:
            p.grad = f.clone()
        else:
            p.grad.copy_(f)


class FlatGradients:
    """Gradient all-reduce without a wrapper module: every parameter's ``.grad`` becomes a VIEW (same sizes and strides as the
    parameter, so fused optimizers see matching layouts) of ONE flat buffer that lives as long as this object; after
    ``loss.backward()`` a single ``reduce()`` averages the whole buffer across ranks in place -- no per-parameter hooks, no
    bucket bookkeeping, no copies.  ``zero()`` replaces ``optimizer.zero_grad()`` (one memset; ``set_to_none`` would drop the views).

    Measured against DistributedDataParallel on 2 B200s (config 2, 19.9 MB of gradients): see DESIGN.md section 5.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], world: int, group=None):
        self.ps = [p for p in params if p.requires_grad]
        self.world, self.group = world, group
        # every view starts on a 256-byte boundary so that the backward's accumulate kernels keep their vector accesses (with
        # optim.FlatAdam, where the PARAMETERS are such views too, unaligned views cost 0.35 ms of the 34.5 ms step); pads stay zero
        pad = lambda k: (k + 63) // 64 * 64      # noqa: E731
        n = sum(pad(p.numel()) for p in self.ps)
        ref = self.ps[0]
        self.flat = torch.zeros(n, device=ref.device, dtype=ref.dtype)
        off = 0
        for p in self.ps:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            if dense:
                p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)
            else:
                p.grad = self.flat[off:off + p.numel()].view(p.size())
            off += pad(p.numel())

    def zero(self) -> None:
        self.flat.zero_()

    def reduce(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)


def gather_labels(local_labels: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Convenience (NOT on the data path): assemble per-rank label blocks [R_local, ...] into [n_total, ...] on every rank."""
    if world == 1:
        return local_labels
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_n = max(e - b for b, e in sizes)
    pad = torch.zeros((max_n,) + tuple(local_labels.shape[1:]), dtype=local_labels.dtype, device=local_labels.device)
    pad[: local_labels.shape[0]] = local_labels
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], 0)
