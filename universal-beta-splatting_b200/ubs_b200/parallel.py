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


class ShardedState:
    """Buffers of the sharded train step: rows are split into `world` shards of `shard_rows` (a multiple of 128);
    rank g owns shard g -- its Adam moments live only there -- and every rank keeps a full copy of the parameters.

      records  [N, stride]                    this rank's parameters, in symmetric (peer-mapped) memory: the owners
                                              of the other shards store updated rows into it over NVLink
      staging  [world, shard_rows, stride]    slot j = rank j's gradient contribution to MY shard, written by rank
                                              j's projection-backward kernel
      exp_avg / exp_avg_sq [shard_rows, stride]   local

    `create` allocates with torch.distributed._symmetric_memory (one process per GPU); `create_local_group` builds
    all ranks' states inside ONE process from ordinary tensors -- the kernels only see device addresses -- which is
    how the single-GPU tests exercise the multi-rank logic bit for bit."""

    def __init__(self, D, N, world, rank, records_buf, staging_buf, peer_records, peer_staging, barrier,
                 rows_buf=None, peer_rows=None):
        from .fused import record_stride

        self.D, self.N, self.world, self.rank = D, N, world, rank
        self.stride = record_stride(D)
        self.shard_rows = shard_rows_for(N, world)
        self._records_buf, self._staging_buf = records_buf, staging_buf
        self.records = records_buf[:N * self.stride].view(N, self.stride)
        self.staging = staging_buf.view(world, self.shard_rows, self.stride)
        self.peer_records, self.peer_staging = list(peer_records), list(peer_staging)
        self.exp_avg = torch.zeros((self.shard_rows, self.stride), dtype=torch.float32, device=records_buf.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self._barrier = barrier
        self.mc_records = None  # NVLS multicast address of the records buffers (set by create() when supported)
        # pull form (default): this rank's [N, 12] screen-space gradient rows live in peer-mapped memory and the shard
        # owners read them from there; "scatter": gradient-record tiles are pushed into the owners' staging buffers
        self._rows_buf = rows_buf
        self.rows = None if rows_buf is None else rows_buf[:N * 12].view(1, N, 12)
        self.peer_rows = None if peer_rows is None else list(peer_rows)
        self.exchange = "pull" if rows_buf is not None else "scatter"

    def attach(self, rz):
        """Make the rasteriser write its screen-space gradient rows into this state's peer-mapped buffer."""
        assert rz.C == 1 and rz.N == self.N and rz.grad_rows and self.rows is not None
        rz.v_rows = self.rows

    def barrier(self):
        """Device-side barrier over the ranks, ordered on the current stream (no host block)."""
        self._barrier()

    def my_rows(self):
        b = self.rank * self.shard_rows
        return b, max(0, min(self.shard_rows, self.N - b))

    @staticmethod
    def create(D: int, N: int, group=None, device=None) -> "ShardedState":
        import torch.distributed._symmetric_memory as symm_mem

        from .fused import record_stride

        group = dist.group.WORLD if group is None else group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        S, stride = shard_rows_for(N, world), record_stride(D)
        rec = symm_mem.empty((world * S * stride,), dtype=torch.float32, device=device)
        stg = symm_mem.empty((world * S * stride,), dtype=torch.float32, device=device)
        rows = symm_mem.empty((world * S * 12,), dtype=torch.float32, device=device)
        rec.zero_()
        stg.zero_()  # the padding rows of a shard are never written and must read as zero gradient
        rows.zero_()
        h_rec, h_stg = symm_mem.rendezvous(rec, group), symm_mem.rendezvous(stg, group)
        h_rows = symm_mem.rendezvous(rows, group)
        st = ShardedState(D, N, world, rank, rec, stg, h_rec.buffer_ptrs, h_stg.buffer_ptrs,
                          lambda: h_rec.barrier(channel=0), rows, h_rows.buffer_ptrs)
        st._handles = (h_rec, h_stg, h_rows)
        st._group = group
        mc = int(h_rec.multicast_ptr or 0)
        st.mc_records = mc or None
        st.barrier()
        return st

    @staticmethod
    def create_local_group(D: int, N: int, world: int, device="cuda"):
        from .fused import record_stride

        S, stride = shard_rows_for(N, world), record_stride(D)
        recs = [torch.zeros((world * S * stride,), dtype=torch.float32, device=device) for _ in range(world)]
        stgs = [torch.zeros((world * S * stride,), dtype=torch.float32, device=device) for _ in range(world)]
        rows = [torch.zeros((world * S * 12,), dtype=torch.float32, device=device) for _ in range(world)]
        return [ShardedState(D, N, world, r, recs[r], stgs[r], [t.data_ptr() for t in recs],
                             [t.data_ptr() for t in stgs], lambda: None, rows[r], [t.data_ptr() for t in rows])
                for r in range(world)]


def shard_rows_for(N: int, world: int, align: int = 128) -> int:
    per = -(-max(N, 1) // world)
    return -(-per // align) * align


@torch.no_grad()
def sharded_backward_scatter(rz, st: ShardedState, viewmats, Ks, cam_pos, timestamps):
    """Projection backward of the most recent composite_backward(): gradient tiles into the owners' staging slots."""
    import ctypes

    from ._lib import check, ptr

    assert rz.C == 1 and rz.N == st.N
    arr = (ctypes.c_void_p * st.world)(*st.peer_staging)
    check(rz.lib.ubs_fused_project_bwd_scatter(
        st.N, st.D, ptr(st.records), ptr(viewmats), ptr(Ks), ptr(cam_pos), ptr(timestamps), rz.W, rz.H, rz.eps2d,
        1 if rz.aa else 0, ptr(rz.radii), ptr(rz.conics), *rz.grad_args(), st.world, st.rank, st.shard_rows,
        ctypes.cast(arr, ctypes.c_void_p), ptr(rz.status), torch.cuda.current_stream().cuda_stream),
        "ubs_fused_project_bwd_scatter")


@torch.no_grad()
def gather_cameras(st: ShardedState, viewmats, Ks, cam_pos, timestamps, group=None):
    """All ranks' camera blocks ([world,4,4], [world,3,3], [world,3], [world]) from this rank's ([1,...] each): one small
    NCCL all-gather, stream-ordered (no host block)."""
    blk = torch.zeros((32,), dtype=torch.float32, device=viewmats.device)
    blk[0:16] = viewmats.reshape(-1)
    blk[16:25] = Ks.reshape(-1)
    blk[25:28] = cam_pos.reshape(-1)
    if timestamps is not None:
        blk[28] = timestamps.reshape(-1)[0]
    flat = torch.empty((st.world * 32,), dtype=torch.float32, device=viewmats.device)
    dist.all_gather_into_tensor(flat, blk, group=group)  # concatenation form: accepted by NCCL and gloo alike
    out = flat.view(st.world, 32)
    return (out[:, 0:16].reshape(st.world, 4, 4).contiguous(), out[:, 16:25].reshape(st.world, 3, 3).contiguous(),
            out[:, 25:28].contiguous(), out[:, 28].contiguous() if timestamps is not None else None)


@torch.no_grad()
def sharded_pull_update(rz, st: ShardedState, adam, viewmats_all, Ks_all, cam_pos_all, timestamps_all,
                        opacity_reg: float = 0.0, scale_reg: float = 0.0, advance: bool = True):
    """Owner side of the pull form, after a barrier behind every rank's composite_backward(): projection backward over
    the `world` views for this rank's shard (the views' gradient rows are read from the ranks' peer-mapped buffers),
    Adam on the shard, new parameters into every rank's records.  *_all: the cameras of all ranks, rank order."""
    import ctypes

    from ._lib import check, ptr

    assert viewmats_all.shape == (st.world, 4, 4) and Ks_all.shape == (st.world, 3, 3) and cam_pos_all.shape == (st.world, 3)
    assert all(t is None or (t.is_contiguous() and t.dtype == torch.float32)
               for t in (viewmats_all, Ks_all, cam_pos_all, timestamps_all))
    if advance:
        adam.step_count += 1
    cols = (ctypes.c_double * st.stride)(*adam.lr_columns())
    recs = (ctypes.c_void_p * st.world)(*st.peer_records)
    rows = (ctypes.c_void_p * st.world)(*st.peer_rows)
    check(rz.lib.ubs_fused_project_bwd_adam_pull(
        st.N, st.D, st.world, st.rank, st.shard_rows, ctypes.cast(recs, ctypes.c_void_p),
        ctypes.cast(rows, ctypes.c_void_p), ptr(viewmats_all), ptr(Ks_all), ptr(cam_pos_all), ptr(timestamps_all),
        rz.W, rz.H, rz.eps2d, 1 if rz.aa else 0, ptr(st.exp_avg), ptr(st.exp_avg_sq),
        ctypes.cast(cols, ctypes.c_void_p), adam.betas[0], adam.betas[1], adam.eps, adam.step_count,
        float(opacity_reg), float(scale_reg), torch.cuda.current_stream().cuda_stream),
        "ubs_fused_project_bwd_adam_pull")


@torch.no_grad()
def sharded_reduce_adam_gather(st: ShardedState, adam, opacity_reg: float = 0.0, scale_reg: float = 0.0,
                               advance: bool = True, use_multicast: bool = False):
    """Owner side: sum the staging slots, Adam on the shard, new parameters into every rank's records.
    `adam` supplies learning rates, betas, eps and the step count (a training.PackedAdam; its own moment buffers
    are not used -- the shard's live in `st`).  use_multicast: broadcast the new rows with one NVLS multimem.st per
    element instead of `world` peer stores; measured SLOWER on B200 (0.63 vs 0.35 ms at 2 GPUs, 0.61 vs 0.57 ms at 8:
    16-byte system-scope multicast stores do not stream), so it is off by default."""
    import ctypes

    from ._lib import check, ptr

    if advance:
        adam.step_count += 1
    cols = (ctypes.c_double * st.stride)(*adam.lr_columns())
    arr = (ctypes.c_void_p * st.world)(*st.peer_records)
    check(adam.lib.ubs_reduce_adam_gather(
        st.N, st.D, st.world, st.rank, st.shard_rows, ptr(st.staging), ptr(st.exp_avg), ptr(st.exp_avg_sq),
        ctypes.cast(arr, ctypes.c_void_p), st.mc_records if use_multicast else None,
        ctypes.cast(cols, ctypes.c_void_p), adam.betas[0], adam.betas[1], adam.eps,
        adam.step_count, float(opacity_reg), float(scale_reg), torch.cuda.current_stream().cuda_stream),
        "ubs_reduce_adam_gather")


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
