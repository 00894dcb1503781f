"""The train step around the render, on the packed records.

    render (fused.FusedRasterizer) -> l1_ssim_loss -> backward -> [NCCL all-reduce] -> PackedAdam.step

mirrors one iteration of train.py:100-171 of the reference: `l1_loss` + `fused_ssim` (train.py:118-121), the
opacity / scale regularisers (:122-124), `total_loss.backward()` (:128), `optimizer.step()` over the seven Adam
groups of scene/beta_model.py:239-268 (:169) and the deterministic part of the MCMC relocation (:150-152).
Every arithmetic step is a kernel of libubs_b200.so; there is no torch fallback.
"""
import ctypes
import math
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr
from .fused import FusedRasterizer, record_slices, record_stride

# learning rates of arguments/__init__.py:84-93 (position_lr_init is scaled by the scene extent by the caller)
DEFAULT_LR = {"xyz": 0.00016, "mean": 0.001, "rgb": 0.001, "opacity": 0.05, "beta": 0.001, "scale": 0.005,
              "l_triangle": 0.001}


def expon_lr(step: int, lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
             max_steps: int = 1000000) -> float:
    """Log-linear learning-rate decay (utils/general_utils.py:34-67, used for the xyz group only)."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    delay = 1.0
    if lr_delay_steps > 0:
        delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    t = min(max(step / max_steps, 0), 1)
    return delay * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


def _strides4(t: Tensor, layout: str):
    """(sn, sc, sy, sx) element strides of a 4-D image tensor given as 'NHWC' or 'NCHW'."""
    assert t.dim() == 4 and t.dtype == torch.float32 and t.is_cuda
    s = t.stride()
    if layout == "NHWC":
        return s[0], s[3], s[1], s[2]
    if layout == "NCHW":
        return s[0], s[1], s[2], s[3]
    raise ValueError(layout)


class _LossWorkspace:
    """Per-shape scratch of the loss kernels, cached per device AND stream (two streams computing the loss of the
    same image shape concurrently must not share scratch)."""

    _cache: Dict[tuple, Tensor] = {}

    @classmethod
    def get(cls, lib, C, ch, H, W, device) -> Tensor:
        key = (C, ch, H, W, str(device), torch.cuda.current_stream(device).cuda_stream)
        ws = cls._cache.get(key)
        if ws is None:
            n = lib.ubs_l1_ssim_workspace_bytes(C, ch, H, W)
            ws = torch.empty((n,), dtype=torch.uint8, device=device)
            cls._cache[key] = ws
        return ws


@torch.no_grad()
def l1_ssim_loss_fwd_bwd(img: Tensor, gt: Tensor, lambda_dssim: float = 0.2, grad_scale: float = 1.0,
                         img_layout: str = "NHWC", gt_layout: str = "NCHW", want_grad: bool = True,
                         v_img: Optional[Tensor] = None, loss_out: Optional[Tensor] = None):
    """Loss values and (optionally) the image gradient in one call.

    img: rendered image(s), [C,H,W,ch] ('NHWC', what the rasteriser writes) or [C,ch,H,W]; gt likewise.
    Returns (loss_out [3] = (L1, SSIM, loss) on the device, v_img with img's shape and strides or None);
    v_img = d(grad_scale * loss) / d img."""
    lib = _lib.load()
    isn, isc, isy, isx = _strides4(img, img_layout)
    gsn, gsc, gsy, gsx = _strides4(gt, gt_layout)
    if img_layout == "NHWC":
        C, H, W, ch = img.shape
    else:
        C, ch, H, W = img.shape
    g_shape = (gt.shape[0], gt.shape[3], gt.shape[1], gt.shape[2]) if gt_layout == "NHWC" else tuple(gt.shape)
    assert g_shape == (C, ch, H, W), "img and gt disagree: %s vs %s" % (tuple(img.shape), tuple(gt.shape))
    ws = _LossWorkspace.get(lib, C, ch, H, W, img.device)
    if loss_out is None:
        loss_out = torch.empty((3,), dtype=torch.float32, device=img.device)
    if want_grad and v_img is None:
        v_img = torch.empty_strided(img.shape, img.stride(), dtype=torch.float32, device=img.device)
    if v_img is not None:
        assert v_img.shape == img.shape and v_img.stride() == img.stride()
    check(lib.ubs_l1_ssim_loss(C, ch, H, W, ptr(img), isn, isc, isy, isx, ptr(gt), gsn, gsc, gsy, gsx,
                               float(lambda_dssim), float(grad_scale), ptr(loss_out),
                               ptr(v_img) if want_grad else None, ptr(ws), ws.numel(),
                               torch.cuda.current_stream().cuda_stream), "ubs_l1_ssim_loss")
    return loss_out, (v_img if want_grad else None)


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt, lambda_dssim, img_layout, gt_layout):
        loss_out, v_img = l1_ssim_loss_fwd_bwd(img.detach(), gt, lambda_dssim, 1.0, img_layout, gt_layout,
                                               want_grad=img.requires_grad)
        ctx.save_for_backward(v_img)
        return loss_out[2].clone()

    @staticmethod
    def backward(ctx, v_loss):
        (v_img,) = ctx.saved_tensors
        return (None if v_img is None else v_img * v_loss), None, None, None, None


def l1_ssim_loss(img: Tensor, gt: Tensor, lambda_dssim: float = 0.2, img_layout: str = "NCHW",
                 gt_layout: str = "NCHW") -> Tensor:
    """Differentiable scalar (1 - lambda) * l1_loss(img, gt) + lambda * (1 - ssim(img, gt)), the expression of
    train.py:118-121, for [C,ch,H,W] (default, the reference's image layout; a [ch,H,W] image is unsqueezed) or
    [C,H,W,ch] tensors.  Gradient flows to `img` only."""
    if img.dim() == 3:
        img, gt = img.unsqueeze(0), gt.unsqueeze(0)
    return _L1SSIM.apply(img, gt, lambda_dssim, img_layout, gt_layout)


class PackedAdam:
    """torch.optim.Adam(lr per group, betas=(0.9, 0.999), eps=1e-15) over the packed record buffer.

    State: exp_avg / exp_avg_sq [N, stride] and the step count, as torch keeps them per parameter
    (scene/beta_model.py:239-268).  `lr` maps the reference's group names to learning rates; `set_lr("xyz", ...)`
    is what `update_learning_rate` (beta_model.py:278-284) does every iteration."""

    GROUPS = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")

    def __init__(self, D: int, N: int, lr: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps: float = 1e-15,
                 device="cuda", allocate_moments: bool = True):
        self.lib = _lib.load()
        self.D, self.N = D, N
        self.stride = record_stride(D)
        self.lr = dict(DEFAULT_LR if lr is None else lr)
        assert set(self.lr) == set(self.GROUPS), "one learning rate per parameter group: %s" % (self.GROUPS,)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        # allocate_moments=False: hyper-parameters and step count only (the sharded step keeps each rank's moment
        # shard in parallel.ShardedState)
        rows = N if allocate_moments else 0
        self.exp_avg = torch.zeros((rows, self.stride), dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros((rows, self.stride), dtype=torch.float32, device=device)

    def set_lr(self, group: str, value: float):
        assert group in self.lr
        self.lr[group] = float(value)

    def lr_columns(self):
        cols = [0.0] * self.stride
        for name, sl in record_slices(self.D).items():
            for c in range(sl.start, sl.stop):
                cols[c] = self.lr[name]
        return cols

    @torch.no_grad()
    def step(self, records: Tensor, grads: Tensor, opacity_reg: float = 0.0, scale_reg: float = 0.0, rows=None,
             advance: bool = True):
        """In-place update of `records` (and the moments) from `grads` ([N, stride], e.g. FusedRasterizer.backward's
        output).  opacity_reg / scale_reg: coefficients of the regularisers of train.py:122-124 (0 = off).
        rows = (begin, count) restricts the update to a row range (one optimiser step applied chunk by chunk:
        pass advance=False for every chunk after the first)."""
        N = records.shape[0]
        assert records.shape == (N, self.stride) and grads.shape == records.shape == self.exp_avg.shape
        assert records.is_cuda and records.is_contiguous() and grads.is_contiguous()
        if advance:
            self.step_count += 1
        begin, count = (0, N) if rows is None else rows
        cols = (ctypes.c_double * self.stride)(*self.lr_columns())
        check(self.lib.ubs_adam_step(N, self.D, begin, count, ptr(records), ptr(grads), ptr(self.exp_avg),
                                     ptr(self.exp_avg_sq), ctypes.cast(cols, ctypes.c_void_p), self.betas[0],
                                     self.betas[1], self.eps, self.step_count, float(opacity_reg), float(scale_reg),
                                     torch.cuda.current_stream().cuda_stream), "ubs_adam_step")

    def grow(self, n_new: int):
        """Zero moments for `n_new` appended rows (cat_tensors_to_optimizer, beta_model.py:446-474)."""
        z = torch.zeros((n_new, self.stride), dtype=torch.float32, device=self.exp_avg.device)
        self.exp_avg = torch.cat((self.exp_avg, z))
        self.exp_avg_sq = torch.cat((self.exp_avg_sq, z))
        self.N += n_new


@torch.no_grad()
def mcmc_relocate(records: Tensor, D: int, dst_idx: Tensor, src_idx: Tensor, adam: Optional[PackedAdam] = None,
                  sharded=None):
    """Rows dst_idx <- rows src_idx with the multiplicity-rescaled opacity; sources take the same opacity and lose
    their Adam moments.  With dst = dead primitives and src = torch.multinomial samples of the alive ones this is
    relocate_gs (scene/beta_model.py:575-620).  Moments: `adam`'s whole-buffer ones, or -- sharded step -- the shard
    of a parallel.ShardedState (every rank calls this with the same indices and resets the sources it owns)."""
    lib = _lib.load()
    N, K = records.shape[0], dst_idx.numel()
    assert src_idx.numel() == K and dst_idx.dtype == torch.int64 and src_idx.dtype == torch.int64
    if K == 0:
        return
    m, v, m_begin, m_count = None, None, 0, N
    if sharded is not None:
        m, v = sharded.exp_avg, sharded.exp_avg_sq
        m_begin, m_count = sharded.rank * sharded.shard_rows, sharded.shard_rows
    elif adam is not None and adam.exp_avg.shape[0] == N:
        m, v = adam.exp_avg, adam.exp_avg_sq
    counts = torch.empty((N,), dtype=torch.int32, device=records.device)
    # bound to locals: a .contiguous() copy must stay alive until after the launch (a temporary's block would be
    # handed to the next same-sized allocation, and both pointers would name the same memory)
    dst_c, src_c = dst_idx.contiguous(), src_idx.contiguous()
    check(lib.ubs_mcmc_relocate(N, D, ptr(records), ptr(m), ptr(v), m_begin, m_count, K, ptr(dst_c), ptr(src_c),
                                ptr(counts), torch.cuda.current_stream().cuda_stream),
          "ubs_mcmc_relocate")


@torch.no_grad()
def mcmc_add(records: Tensor, D: int, src_idx: Tensor, adam: Optional[PackedAdam] = None) -> Tensor:
    """add_new_gs (scene/beta_model.py:622-657) given the sampled sources: returns the grown record buffer whose
    appended rows are the rescaled copies; the moments grow with zeros."""
    K = src_idx.numel()
    if K == 0:
        return records
    N = records.shape[0]
    grown = torch.cat((records, torch.empty((K, records.shape[1]), dtype=records.dtype, device=records.device)))
    if adam is not None:
        adam.grow(K)
    dst = torch.arange(N, N + K, dtype=torch.int64, device=records.device)
    mcmc_relocate(grown, D, dst, src_idx, adam)
    return grown


@torch.no_grad()
def sgld_noise(records: Tensor, D: int, noise_lr: float, xyz_lr: float, noise: Optional[Tensor] = None,
               generator=None) -> Tensor:
    """The position noise of the MCMC step (train.py:156-163), in place on the xyz columns:
    xyz += get_xyz_covariance @ (randn * (1 - opacity)^100 * noise_lr * xyz_lr).  `noise` ([N,3], N(0,1)) defaults to a
    torch.randn draw on the records' device (torch's RNG stream defines it, as in the reference).  Returns the noise."""
    lib = _lib.load()
    N = records.shape[0]
    if noise is None:
        noise = torch.randn((N, 3), dtype=torch.float32, device=records.device, generator=generator)
    assert noise.shape == (N, 3) and noise.dtype == torch.float32 and noise.device == records.device
    noise = noise.contiguous()
    check(lib.ubs_sgld_noise(N, D, ptr(records), ptr(noise), float(noise_lr), float(xyz_lr),
                             torch.cuda.current_stream().cuda_stream), "ubs_sgld_noise")
    return noise


def sample_alive(probs: Tensor, num: int, alive_indices: Optional[Tensor] = None, generator=None) -> Tensor:
    """_sample_alives (scene/beta_model.py:567-573) without the bincount: the multinomial draw itself is torch's
    (its result is defined by torch's RNG stream)."""
    probs = probs / (probs.sum() + torch.finfo(torch.float32).eps)
    idx = torch.multinomial(probs, num, replacement=True, generator=generator)
    return idx if alive_indices is None else alive_indices[idx]


class TrainStep:
    """One full training iteration on resident buffers, zero host syncs:
    forward -> L1+SSIM loss and its image gradient -> backward -> (all-reduce) -> Adam.

    The rasteriser's camera count C is the per-GPU batch: train.py:111-128 renders `batch_size` views, averages their
    losses, calls backward() once and steps the optimiser once -- here the C views are rendered by ONE launch sequence,
    the loss kernel averages over the C images, the projection backward sums the per-camera gradients in registers,
    and there is one update.  With `world` ranks the effective batch is C * world (pass it as `batch_size`)."""

    def __init__(self, rz: FusedRasterizer, adam: PackedAdam, lambda_dssim: float = 0.2, world: int = 1, group=None,
                 fuse_adam: bool = True, n_chunks: int = 4, sharded=None):
        assert sharded is None or rz.C == 1, "the sharded step scatters one view per rank"
        # row-range pipelining of the projection backward against the all-reduce needs [1, N] screen-space arrays
        self.n_chunks = n_chunks if rz.C == 1 else 1
        # sharded: a parallel.ShardedState -- gradient tiles go over NVLink into the owner rank's staging buffer
        # straight from the projection-backward kernel, the owner reduces + applies Adam to its 1/world of the rows
        # and stores the new parameters into every rank's records; `records` passed to step() must be sharded.records
        self.sharded = sharded
        self.rz, self.adam, self.lam, self.world, self.group = rz, adam, lambda_dssim, world, group
        # nothing to sum across ranks before the update on one GPU (the reference's default set-up): Adam rides in the
        # projection-backward kernel and the gradient records never reach HBM
        self.fuse_adam = fuse_adam
        dev = rz.device
        self.v_rc = torch.empty_like(rz.render_colors)
        self.v_ra = torch.zeros_like(rz.render_alphas)  # the loss does not depend on alpha
        self.loss_out = torch.empty((3,), dtype=torch.float32, device=dev)
        self.v_records = None
        # sharded pull form: the cameras of ALL ranks for the coming step, ([world,4,4], [world,3,3], [world,3], [world] or
        # None), when the caller knows them; otherwise step() all-gathers them
        self.all_cameras = None

    @torch.no_grad()
    def step(self, records: Tensor, viewmats, Ks, cam_pos, timestamps, backgrounds, gt: Tensor, gt_layout="NCHW",
             opacity_reg: float = 0.0, scale_reg: float = 0.0, batch_size: Optional[int] = None,
             apply_update: bool = True) -> Tensor:
        """gt: [C,3,H,W] (or [3,H,W] for C = 1).  batch_size: views per optimiser step over all ranks (default
        C * world); each rank's loss is the mean over its C views scaled by C / batch_size.
        apply_update=False: forward, loss and backward only -- the gradient records are left in `self.v_records`, the
        parameters, the moments and the step count stay untouched.  (On its densification iterations the reference
        swaps fresh nn.Parameters into the optimiser BEFORE optimizer.step(), so that step finds no gradients and is
        skipped, train.py:150-169 / scene/beta_model.py:512-546: a caller reproducing that ordering passes False.)"""
        from . import parallel

        rz = self.rz
        C = rz.C
        batch_size = C * self.world if batch_size is None else batch_size
        fused_update = self.fuse_adam and self.world == 1 and self.sharded is None and apply_update
        if not fused_update and self.sharded is None and (self.v_records is None or
                                                          self.v_records.shape != records.shape):
            self.v_records = torch.empty_like(records)
        rc, _ = rz.forward(records, viewmats, Ks, cam_pos, timestamps, backgrounds)
        if gt.dim() == 3:
            gt = gt.unsqueeze(0)
        with rz._stage("l1_ssim_loss"):
            # the kernel's loss is the mean over the C images; d(C / batch_size * that) reaches the renderer
            l1_ssim_loss_fwd_bwd(rc, gt, self.lam, float(C) / batch_size, "NHWC", gt_layout, True, self.v_rc,
                                 self.loss_out)
        if fused_update:
            rz.backward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, self.v_rc, self.v_ra, None, self.adam,
                        opacity_reg, scale_reg)
            return self.loss_out
        if not apply_update:
            assert self.sharded is None, "apply_update=False needs a local gradient buffer (not the sharded step)"
            rz.backward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, self.v_rc, self.v_ra, self.v_records)
            if self.world > 1:
                parallel.allreduce_gradients(self.v_records, self.world, None, self.group)
            return self.loss_out
        if self.sharded is not None:
            st = self.sharded
            assert records.data_ptr() == st.records.data_ptr(), "step() must be given ShardedState.records"
            if st.exchange == "pull":
                # the 48-byte screen-space gradient rows are what crosses NVLink: the owner of a shard reads all views'
                # rows of its primitives, runs projection backward + Adam for them and stores the new rows everywhere
                if rz.v_rows is not st.rows:
                    st.attach(rz)
                cams = (self.all_cameras if self.all_cameras is not None else
                        parallel.gather_cameras(st, viewmats, Ks, cam_pos, timestamps, self.group))
                rz.composite_backward(backgrounds, self.v_rc, self.v_ra)
                with rz._stage("barrier"):
                    st.barrier()  # every rank's rows are complete
                with rz._stage("pull_bwd_adam_gather"):
                    parallel.sharded_pull_update(rz, st, self.adam, *cams, opacity_reg, scale_reg)
                with rz._stage("barrier"):
                    st.barrier()  # every shard's new parameters have landed in my records; my rows have been read
                return self.loss_out
            rz.composite_backward(backgrounds, self.v_rc, self.v_ra)
            with rz._stage("bwd_scatter"):
                parallel.sharded_backward_scatter(rz, st, viewmats, Ks, cam_pos, timestamps)
            with rz._stage("barrier"):
                st.barrier()  # every rank's contribution to my shard has landed
            with rz._stage("reduce_adam_gather"):
                parallel.sharded_reduce_adam_gather(st, self.adam, opacity_reg, scale_reg)
            with rz._stage("barrier"):
                st.barrier()  # every shard's new parameters have landed in my records
            return self.loss_out
        # world > 1, or the unfused single-GPU form: the projection backward runs chunk by chunk; chunk k's gradient
        # all-reduce (NCCL, its own stream) overlaps the backward of chunk k+1 and the Adam update of chunk k-1
        adam, first = self.adam, [True]

        def after_reduce(begin, count):
            adam.step(records, self.v_records, opacity_reg, scale_reg, rows=(begin, count), advance=first[0])
            first[0] = False

        with rz._stage("bwd_allreduce_adam"):
            parallel.pipelined_backward(rz, records, viewmats, Ks, cam_pos, timestamps, backgrounds, self.v_rc,
                                        self.v_ra, self.v_records, self.world, self.group, self.n_chunks, after_reduce)
        return self.loss_out
