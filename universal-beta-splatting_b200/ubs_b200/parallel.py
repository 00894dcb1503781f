"""Multi-GPU use of the rasteriser on one NVSwitch box: one process per GPU (torchrun), torch.distributed for the
plumbing.  The reference has no multi-GPU path at all; this is what BASELINE.json's north star adds.

  * camera-parallel rendering: cameras are independent units -> round-robin sharding, NO collective on the data
    path (`shard_cameras`); every rank holds the full (read-only) parameter records.
  * data-parallel training over views: every rank runs forward+backward on its views into ONE flat packed
    gradient buffer [N, stride] (the layout the fused backward kernel writes), then a single
    all_reduce(sum) over NCCL/NVLink and a 1/B scale (train.py:127 divides the loss by the batch size).
"""
import os
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def init_process_group(backend: str = "nccl", **kw):
    """torch.distributed.init_process_group with NCCL's internal stream at high priority (must be set before the
    group is created): the chunk-pipelined all-reduce of `pipelined_backward` shares the GPU with full-machine
    grids and otherwise only starts when they drain."""
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    return dist.init_process_group(backend, **kw)


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


def row_chunks(N: int, n_chunks: int, align: int = 128):
    """[(begin, count)] covering [0, N) in at most n_chunks pieces whose starts are multiples of `align` (the
    projection-backward kernel's CTA span, so chunked and unchunked runs launch identical CTAs)."""
    per = max(align, -(-N // max(n_chunks, 1)))
    per = -(-per // align) * align
    return [(b, min(per, N - b)) for b in range(0, N, per)]


@torch.no_grad()
def pipelined_backward(rz, records, viewmats, Ks, cam_pos, timestamps, backgrounds, v_render_colors,
                       v_render_alphas, v_records: Tensor, world: int, group=None, n_chunks: int = 4,
                       after_reduce=None, batch_size: Optional[int] = None):
    """Backward of one view per rank with the gradient all-reduce overlapped: the compositing backward runs once,
    then for each chunk of primitives the projection backward is followed by an asynchronous all_reduce(sum) of that
    chunk's gradient rows -- NCCL runs it on its own stream, so it overlaps the next chunk's kernel -- and
    `after_reduce(begin, count)` (e.g. the Adam update of those rows) is queued behind the chunk's reduction."""
    rz.composite_backward(backgrounds, v_render_colors, v_render_alphas)
    chunks = row_chunks(rz.N, n_chunks if world > 1 else 1)
    works = []
    for begin, count in chunks:
        rz.project_backward_rows(records, viewmats, Ks, cam_pos, timestamps, v_records, begin, count)
        rows = v_records[begin:begin + count]
        works.append(dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group, async_op=True) if world > 1 else None)
    for (begin, count), work in zip(chunks, works):
        if work is not None:
            work.wait()  # orders the current stream behind this chunk's reduction (no host block on CUDA)
        if batch_size is not None and batch_size != 1:
            v_records[begin:begin + count].mul_(1.0 / batch_size)
        if after_reduce is not None:
            after_reduce(begin, count)
    return v_records


class DataParallelTrainer:
    """forward + backward of one view per rank with the (chunk-pipelined) gradient all-reduce.
    `rz` is a fused.FusedRasterizer."""

    def __init__(self, rz, world: int, group=None, n_chunks: int = 4):
        self.rz, self.world, self.group, self.n_chunks = rz, world, group, n_chunks

    @torch.no_grad()
    def step(self, records, viewmats, Ks, cam_pos, timestamps, backgrounds, v_render_colors, v_render_alphas,
             v_records: Tensor, batch_size: Optional[int] = None) -> Tensor:
        args = (records, viewmats, Ks, cam_pos, timestamps, backgrounds)
        self.rz.forward(*args)
        if self.world == 1:
            self.rz.backward(*args, v_render_colors, v_render_alphas, v_records)
            allreduce_gradients(v_records, self.world, batch_size, self.group)
            return v_records
        return pipelined_backward(self.rz, *args, v_render_colors, v_render_alphas, v_records, self.world, self.group,
                                  self.n_chunks, None, batch_size)
