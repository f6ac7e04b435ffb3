"""Data parallelism: one process per GPU, one NCCL all-reduce of the flat gradient buffer per step.

Replaces `keras.utils.multi_gpu_model` (utils.py:209-211): the reference slices the batch over towers inside one TF
graph and sums the tower gradients implicitly through shared variables; BatchNorm statistics stay per tower.  Here
each rank holds a replica and a batch shard, BatchNorm stays per replica (same semantics), and the 2.11 M fp32
gradients (8.45 MB) are summed over NVLink/NVSwitch in two buckets INSIDE the captured step: the suffix of the flat
buffer (blocks 13..16 + ASPP + head, 76 % of the parameters) as soon as the backward pass has produced it, overlapped
with the remaining backward kernels, and the 2 MB prefix at the end; Adam divides by world_size (dlb_adam_step
grad_mult).  The path has no other exchange step.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_process_group_from_env(backend: str = "nccl"):
    """torchrun-style env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def make_data_parallel(model, group=None):
    """Attach the gradient all-reduce to the model's engine (no-op for world_size 1)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return model
    e = model.engine
    e.world_size = dist.get_world_size(group)

    def hook(flat_grads: torch.Tensor, async_op: bool = False):
        """sum a (slice of the) flat gradient buffer over the ranks; async_op=True returns the work handle so the
        engine can launch the next backward kernels before the collective has finished"""
        return dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    e.grad_hook = hook
    e._graphs.clear()
    # replicas must start identical: broadcast rank 0's parameters and BN statistics
    dist.broadcast(e.params, src=0, group=group)
    dist.broadcast(e.stats, src=0, group=group)
    e._weights_dirty = True
    return model


def shard_batch(n_items: int, rank: int, world: int):
    """contiguous shard [lo, hi) of a global batch (the reference's multi_gpu_model slices the batch the same way)."""
    per = n_items // world
    if per * world != n_items:
        raise ValueError(f"global batch {n_items} is not divisible by world size {world}")
    return rank * per, (rank + 1) * per
