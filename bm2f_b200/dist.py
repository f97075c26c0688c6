"""Data-parallel plumbing for the hot path: images are independent, so the batch is
split across ranks with no data-path collective; the only exchange is the DDP-style all-reduce of
the MSDeformAttn projection-weight gradients (4 Linear x n_layers = 1,233,600 fp32 = 4.93 MB for the
6-layer encoder), done as ONE flat bucket.  The reference gets this from Detectron2's
DistributedDataParallel (train_net.py:325-335); nothing in its tree implements it."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch


def shard_batch(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `total` images over `world` ranks: returns (first_image, count).
    The first total % world ranks take one extra image."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def trainable_parameters(module: torch.nn.Module) -> List[torch.nn.Parameter]:
    """EVERY parameter of an encoder / pixel decoder that takes a gradient, in registration order: what a data-parallel
    training step has to exchange when the whole mirror (FFN, LayerNorm, level_embed, input_proj, FPN convolutions) is
    trained, not only the four MSDeformAttn projections (use with GradBucket, or wrap the module in DDP)."""
    return [p for p in module.parameters() if p.requires_grad]


def projection_parameters(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    """Parameters of the hot path that need gradient exchange, in a deterministic order:
    for every MSDeformAttn-like module its sampling_offsets, attention_weights, value_proj and
    output_proj (weight, bias)."""
    out = []
    for m in modules:
        for name in ("sampling_offsets", "attention_weights", "value_proj", "output_proj"):
            lin = getattr(m, name)
            out += [lin.weight, lin.bias]
    return out


class GradBucket:
    """One flat fp32 bucket for a list of parameters' gradients; all_reduce(sum) then scale by 1/world (DDP semantics).
    With `projection_parameters` it is the hot path's own exchange (bench.py); a training loop that updates the rest of
    the encoder / decoder mirror must bucket `trainable_parameters(model)` instead, or those gradients are never
    reduced.  pack() / flat / unpack() are separate so that the caller can run the collective on a side stream."""

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = params
        self.group = group
        self.numel = sum(p.numel() for p in params)
        dev = params[0].device if params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in params:                                   # the bucket seen through the parameters' shapes
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def pack(self):
        """gradients -> bucket: one multi-tensor copy (a handful of launches, not one per parameter)"""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        if grads:
            torch._foreach_copy_(self.views, grads)

    def unpack(self):
        for p in self.params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
        if self.params:
            torch._foreach_copy_([p.grad for p in self.params], self.views)

    def all_reduce(self):
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)
        self.unpack()


def max_over_ranks(value: float, device, group=None) -> float:
    """Max of a host scalar over all ranks (multi-GPU timings are reported as the slowest rank)."""
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
