"""Multi-GPU use of the rasteriser on one NVSwitch box: one process per GPU (torchrun), torch.distributed for the
plumbing.  The reference has no multi-GPU path at all; this is what BASELINE.json's north star adds.

  * camera-parallel rendering: cameras are independent units -> round-robin sharding, NO collective on the data
    path (`shard_cameras`); every rank holds the full (read-only) parameter records.
  * data-parallel training over views: every rank runs forward+backward on its views into ONE flat packed
    gradient buffer [N, stride] (the layout the fused backward kernel writes), then a single
    all_reduce(sum) over NCCL/NVLink and a 1/B scale (train.py:127 divides the loss by the batch size).
"""
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def shard_cameras(n_cameras: int, world: int, rank: int) -> List[int]:
    """Indices of the cameras rank `rank` renders: c with c mod world == rank (balanced to within one camera)."""
    assert 0 <= rank < world
    return list(range(rank, n_cameras, world))


def allreduce_gradients(v_records: Tensor, world: int, batch_size: Optional[int] = None, group=None,
                        async_op: bool = False):
    """Sum the packed gradient records over ranks in place (one collective for all 7 parameter tensors) and scale
    by 1/batch_size when given.  Works on NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    work = None
    if world > 1:
        work = dist.all_reduce(v_records, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if batch_size is not None and batch_size != 1:
        if work is not None and async_op:
            work.wait()
            work = None
        v_records.mul_(1.0 / batch_size)
    return work


class DataParallelTrainer:
    """forward + backward of one view per rank, then the gradient all-reduce.  `rz` is a fused.FusedRasterizer."""

    def __init__(self, rz, world: int, group=None):
        self.rz, self.world, self.group = rz, world, group

    @torch.no_grad()
    def step(self, records, viewmats, Ks, cam_pos, timestamps, backgrounds, v_render_colors, v_render_alphas,
             v_records: Tensor, batch_size: Optional[int] = None) -> Tensor:
        args = (records, viewmats, Ks, cam_pos, timestamps, backgrounds)
        self.rz.forward(*args)
        self.rz.backward(*args, v_render_colors, v_render_alphas, v_records)
        allreduce_gradients(v_records, self.world, batch_size, self.group)
        return v_records
