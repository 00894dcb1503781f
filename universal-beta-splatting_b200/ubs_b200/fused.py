"""Fused fast path: packed primitive records + cameras -> image (and back), with zero host synchronisation.

This is the B200-first restructuring of `BetaModel.render` (scene/beta_model.py:660-722): the seven raw parameter
tensors live in ONE packed [N, stride] FP32 buffer (layout in include/ubs_b200.h), one kernel does the
activations + covariance build + conditioning + projection + tile count for all cameras, the pair list is
capacity-bounded so its length never has to visit the host, and all scratch is allocated once and reused.
"""
import math
import warnings
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import UbsError, check, ptr


def record_stride(D: int) -> int:
    return (3 * D + 2 + D * (D - 1) // 2 + 3) // 4 * 4


def record_slices(D: int):
    """Column ranges of the packed record (must match include/ubs_b200.h)."""
    o = 0
    out = {}
    for name, w in (("xyz", 3), ("mean", D - 3), ("rgb", 3), ("opacity", 1), ("beta", D - 2), ("scale", D),
                    ("l_triangle", D * (D - 1) // 2)):
        out[name] = slice(o, o + w)
        o += w
    return out


def pack_records(D, xyz, mean, rgb, opacity, beta, scale, l_triangle) -> Tensor:
    """7 raw BetaModel tensors (scene/beta_model.py:57-63) -> packed [N, stride] records (zero padded)."""
    N = xyz.shape[0]
    rec = torch.zeros((N, record_stride(D)), dtype=torch.float32, device=xyz.device)
    sl = record_slices(D)
    for name, t in (("xyz", xyz), ("mean", mean), ("rgb", rgb), ("opacity", opacity.reshape(N, 1)), ("beta", beta),
                    ("scale", scale), ("l_triangle", l_triangle)):
        rec[:, sl[name]] = t
    return rec


def unpack_records(D, rec: Tensor):
    """Views into the packed buffer, in BetaModel order (xyz, mean, rgb, opacity, beta, scale, l_triangle)."""
    sl = record_slices(D)
    return tuple(rec[:, sl[n]] for n in ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"))


class _StageTimer:
    """Records a CUDA-event pair around a stage on the current stream when stage timing is enabled."""

    def __init__(self, sink, name):
        self.sink, self.name = sink, name

    def __enter__(self):
        if self.sink is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.sink is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.sink.setdefault(self.name, []).append((self.e0, e1))
        return False


class FusedRasterizer:
    """Persistent buffers + the launch sequence fused-project -> emit/sort/offsets -> composite.

    `capacity` bounds the number of (primitive, tile) pairs (default 8 C N; a 3M-primitive 1080p frame has ~2 N).
    A frame that exceeds it is TRUNCATED on the device, and that is never silent:
      * the device word `status[0]` says whether the most recent frame was truncated, `status[1]` counts truncated
        frames; every projection-backward entry point takes `status` as its skip flag, so a truncated frame contributes
        no gradient and the fused Adam update of such a frame is not applied at all;
      * pair count and status of every frame are copied to pinned host memory asynchronously and inspected (without
        blocking) by the following calls: the buffers grow BEFORE they overflow (at 85 % fill) and, if a frame was
        truncated all the same, `on_overflow` decides: "warn" (default) emits a RuntimeWarning, "raise" raises
        UbsError from the next forward(); `truncated_frames` counts them;
      * the autograd wrapper `render()` additionally refuses to differentiate a frame it knows to be truncated.
    `frame_id` numbers the forward() calls; backward() works on the most recent frame only and the autograd wrapper
    raises when asked to differentiate a stale one.
    """

    def __init__(self, D: int, N: int, width: int, height: int, n_cams: int = 1, capacity: Optional[int] = None,
                 tile_size: int = 16, device="cuda", eps2d=0.3, near_plane=0.01, far_plane=1e10, radius_clip=0.0,
                 antialiased=False, sort_mode: str = "bin", on_overflow: str = "warn", grad_rows: bool = True):
        self.lib = _lib.load()
        self.grad_rows = grad_rows  # RGB frames: screen-space gradients as 48-byte rows (False: separate arrays)
        self.v_rows = None          # [C, N, 12], allocated by the first backward (or attached by parallel.ShardedState)
        self.D, self.N, self.W, self.H, self.C = D, N, width, height, n_cams
        self.tile_size = tile_size
        self.tw, self.th = math.ceil(width / tile_size), math.ceil(height / tile_size)
        self.eps2d, self.near, self.far, self.clip, self.aa = eps2d, near_plane, far_plane, radius_clip, antialiased
        self.device = torch.device(device)
        # "bin": tile binning + per-tile segment sort (csrc/bin_sort.cu); "onesweep": emit + global onesweep radix
        # sort (csrc/isect.cu, csrc/radix_sort.cu).  Both give bit-identical lists; binning needs depths >= 0.
        assert sort_mode in ("bin", "onesweep")
        self.sort_mode = sort_mode if near_plane > 0 else "onesweep"
        dev = self.device
        C = n_cams
        f32, i32 = torch.float32, torch.int32
        self.radii = torch.empty((C, N), dtype=i32, device=dev)
        self.means2d = torch.empty((C, N, 2), dtype=f32, device=dev)
        self.depths = torch.empty((C, N), dtype=f32, device=dev)
        self.conics = torch.empty((C, N, 3), dtype=f32, device=dev)
        self.opacities = torch.empty((C, N), dtype=f32, device=dev)
        self.betas = torch.empty((C, N), dtype=f32, device=dev)
        self.colors = torch.empty((C, N, 3), dtype=f32, device=dev)
        # the same screen-space record as one 48-byte row per primitive (what the compositing kernels gather from)
        self.splats = torch.empty((C, N, 12), dtype=f32, device=dev)
        self.tiles_per_gauss = torch.empty((C, N), dtype=i32, device=dev)
        self.n_isects = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.status = torch.zeros((2,), dtype=i32, device=dev)  # [truncated now, truncated frames so far]
        assert on_overflow in ("warn", "raise")
        self.on_overflow = on_overflow
        self.frame_id = 0          # number of forward() calls so far; the frame backward() differentiates
        self.truncated_frames = 0  # as seen by the host (lags the device by a frame or two)
        self.channels = 3          # colour channels of the most recent frame (3 RGB, 4 RGB+depth, 1 depth)
        self._extra_out = {}
        self.offsets = torch.empty((C, self.th, self.tw), dtype=i32, device=dev)
        self.render_colors = torch.empty((C, height, width, 3), dtype=f32, device=dev)
        self.render_alphas = torch.empty((C, height, width, 1), dtype=f32, device=dev)
        self.last_ids = torch.empty((C, height, width), dtype=i32, device=dev)
        self._host_count = torch.zeros((2,), dtype=torch.int64).pin_memory()
        self._host_status = torch.zeros((2,), dtype=torch.int32).pin_memory()
        self._count_event, self._count_frame = None, 0
        self._alloc_pairs(capacity if capacity is not None else max(8 * C * N, 1 << 16))
        self.stage_events = None  # dict stage -> [(start, end)] when enable_stage_timing() was called

    # ---- optional per-stage CUDA-event timing (bench.py's roofline bookkeeping) --------------------------------
    def enable_stage_timing(self, on: bool = True):
        self.stage_events = {} if on else None

    def _stage(self, name: str):
        return _StageTimer(self.stage_events, name)

    def stage_times_ms(self):
        """{stage: (launch count, mean ms)} over everything recorded since enable_stage_timing(); synchronises."""
        torch.cuda.synchronize(self.device)
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1))
                for k, v in (self.stage_events or {}).items()}

    def work_counts(self):
        """Compositing work counters of the most recent forward() (diagnostic kernel):
        dict(E_test, E_acc, E_cull, pairs_staged, pairs, visible)."""
        assert getattr(self, "has_screen_space", True), "work_counts() needs a frame rendered with screen_space=True"
        counts = torch.zeros((8,), dtype=torch.int64, device=self.device)
        check(self.lib.ubs_rasterize_count(
            self.C, ptr(self.n_isects), self.capacity, ptr(self.means2d), ptr(self.conics), ptr(self.opacities),
            ptr(self.betas), self.W, self.H, self.tile_size, ptr(self.offsets), ptr(self.flatten_ids), ptr(counts),
            torch.cuda.current_stream().cuda_stream), "ubs_rasterize_count")
        c = counts.tolist()
        return dict(E_test=c[0], E_acc=c[1], E_cull=c[2], pairs_staged=c[3], E_any=c[4], E_cull4=c[5], E_any4=c[6],
                    E_any8x2=c[7], pairs=self.last_pair_count(), visible=int((self.radii > 0).sum()))

    def _alloc_pairs(self, capacity: int):
        self.capacity = int(capacity)
        dev = self.device
        self.isect_ids = torch.empty((self.capacity,), dtype=torch.int64, device=dev)
        self.flatten_ids = torch.empty((self.capacity,), dtype=torch.int32, device=dev)
        if self.sort_mode == "bin":
            nbytes = self.lib.ubs_isect_bin_workspace_bytes(self.C, self.tw, self.th, self.capacity)
        else:
            nbytes = self.lib.ubs_isect_workspace_bytes(self.C * self.N, self.capacity)
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=dev)

    def _poll_count(self, block: bool = False):
        """Look at the pair count / truncation status an earlier frame left in pinned memory (non-blocking unless
        `block`); grows the pair buffers ahead of need and reports truncated frames (class docstring)."""
        if self._count_event is None:
            return
        if block:
            self._count_event.synchronize()
        elif not self._count_event.query():
            return
        n, n_trunc = int(self._host_count[0]), int(self._host_status[1])
        self._count_event = None
        if n > 0.85 * self.capacity:
            self._alloc_pairs(int(n * 1.5) + 1024)
        if n_trunc > self.truncated_frames:
            new = n_trunc - self.truncated_frames
            self.truncated_frames = n_trunc
            msg = ("%d frame(s) exceeded the pair capacity and were truncated (last pair count %d); their gradients "
                   "were dropped and the buffers have been grown to %d pairs" % (new, n, self.capacity))
            if self.on_overflow == "raise":
                raise UbsError(msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def last_pair_count(self) -> int:
        """Blocking read of the last frame's pair count (diagnostics / tests)."""
        return int(self.n_isects.item())

    def overflowed(self) -> bool:
        """Was the most recent frame truncated?  (Blocking read of the device status word.)"""
        return bool(self.status[0].item())

    def _color_buffer(self, channels: int) -> Tensor:
        if channels == 3:
            return self.render_colors
        buf = self._extra_out.get(channels)
        if buf is None:
            buf = torch.empty((self.C, self.H, self.W, channels), dtype=torch.float32, device=self.device)
            self._extra_out[channels] = buf
        return buf

    @torch.no_grad()
    def forward(self, records: Tensor, viewmats: Tensor, Ks: Tensor, cam_pos: Tensor,
                timestamps: Optional[Tensor] = None, backgrounds: Optional[Tensor] = None,
                prim_mask: Optional[Tensor] = None, out=None, channels: int = 3, activated: bool = False,
                query: Optional[Tensor] = None, screen_space: bool = True):
        """records [N,stride], viewmats [C,4,4], Ks [C,3,3], cam_pos [C,3], timestamps [C] (D=7),
        backgrounds [C,channels] -> (render_colors [C,H,W,channels], render_alphas [C,H,W,1]).  The images land in
        buffers owned by self, or in `out` = (colors, alphas) when given (backward() needs the default alphas).
        channels: 3 = RGB; 4 = RGB + depth ("RGB+D" / "RGB+ED"); 1 = depth ("Depth" / "EDepth" / "Normal") --
        submodules/gsplat/rendering.py:131-142; the depth channel comes out of the same 48-byte splat rows.
        activated / query: see ubs_fused_project_fwd (the drop-in route).
        screen_space=False: render-only frame -- the separate conics / opacities / betas / colors / tiles_per_gauss
        arrays (what backward(), work_counts() and `meta` read) are not written; only with the tile-binning route."""
        lib, s = self.lib, torch.cuda.current_stream().cuda_stream
        C, N, D = self.C, self.N, self.D
        assert records.shape == (N, record_stride(D)) and records.is_cuda and records.dtype == torch.float32
        assert records.is_contiguous()
        assert viewmats.shape == (C, 4, 4) and Ks.shape == (C, 3, 3)
        assert query is not None or cam_pos.shape == (C, 3)
        assert query is None or (query.shape == (N, D - 3) and query.is_contiguous() and query.dtype == torch.float32)
        assert channels in (1, 3, 4)
        assert backgrounds is None or (backgrounds.shape == (C, channels) and backgrounds.is_contiguous())
        self._poll_count()
        self.frame_id += 1
        self.channels = channels
        rc_out, ra_out = (self._color_buffer(channels), self.render_alphas) if out is None else out
        assert rc_out.shape == (C, self.H, self.W, channels) and ra_out.shape == self.render_alphas.shape
        assert rc_out.is_contiguous() and ra_out.is_contiguous() and rc_out.dtype == torch.float32
        mask_u8 = None if prim_mask is None else prim_mask.to(torch.bool).contiguous().view(torch.uint8)
        full = screen_space or self.sort_mode != "bin"
        self.has_screen_space = full
        with self._stage("fused_project_fwd"):
          check(lib.ubs_fused_project_fwd(
            C, N, D, ptr(records), ptr(viewmats), ptr(Ks), ptr(cam_pos), ptr(timestamps), ptr(mask_u8), self.W, self.H,
            self.eps2d, self.near, self.far, self.clip, 1 if self.aa else 0, self.tile_size, self.tw, self.th,
            ptr(self.radii), ptr(self.means2d), ptr(self.depths), ptr(self.conics) if full else None,
            ptr(self.opacities) if full else None, ptr(self.betas) if full else None,
            ptr(self.colors) if full else None, ptr(self.tiles_per_gauss) if full else None, ptr(self.splats),
            ptr(self.workspace) if self.sort_mode == "bin" else None, ptr(self.n_isects), ptr(self.workspace),
            self.workspace.numel(), 1 if activated else 0, ptr(query), s), "ubs_fused_project_fwd")
        with self._stage("isect_emit_sort_offsets"):
          if self.sort_mode == "bin":
            check(lib.ubs_isect_bin_sort(
              C, N, ptr(self.means2d), ptr(self.radii), ptr(self.depths), self.tile_size, self.tw, self.th, 1, None,
              self.capacity, ptr(self.n_isects), ptr(self.isect_ids), ptr(self.flatten_ids), ptr(self.offsets),
              ptr(self.status), ptr(self.workspace), self.workspace.numel(), s), "ubs_isect_bin_sort")
          else:
            check(lib.ubs_isect_emit_sort(
              C, N, ptr(self.means2d), ptr(self.radii), ptr(self.depths), self.tile_size, self.tw, self.th, 1,
              ptr(self.tiles_per_gauss), ptr(self.n_isects), self.capacity, ptr(self.isect_ids),
              ptr(self.flatten_ids), ptr(self.offsets), ptr(self.status), ptr(self.workspace),
              self.workspace.numel(), s), "ubs_isect_emit_sort")
        with self._stage("rasterize_fwd"):
          check(lib.ubs_rasterize_fwd_splats(
            C, N, ptr(self.n_isects), self.capacity, ptr(self.splats), None, ptr(backgrounds), None, channels, self.W,
            self.H, self.tile_size, ptr(self.offsets), ptr(self.flatten_ids), ptr(rc_out), ptr(ra_out),
            ptr(self.last_ids), s), "ubs_rasterize_fwd_splats")
        if self._count_event is None:
            self._host_count[0:1].copy_(self.n_isects, non_blocking=True)
            self._host_status.copy_(self.status, non_blocking=True)
            self._count_event = torch.cuda.Event()
            self._count_event.record()
            self._count_frame = self.frame_id
        return rc_out, ra_out

    def _grad_buffers(self):
        """Screen-space gradient storage of the frame's mode, to be zeroed before the compositing backward.
        RGB frames: one 48-byte row per (camera, primitive), v_rows [C, N, 12] (ubs_rasterize_bwd_rows; layout in
        include/ubs_b200.h).  Depth modes (and grad_rows=False): the separate arrays of ubs_rasterize_bwd_splats,
        which composite_backward() then packs into v_rows (the projection backward reads rows only)."""
        C, N = self.C, self.N
        if getattr(self, "v_rows", None) is None:
            self.v_rows = torch.empty((C, N, 12), dtype=torch.float32, device=self.device)
        if self.channels == 3 and self.grad_rows:
            return self.v_rows
        if getattr(self, "_gflat", None) is None:
            self._gflat = torch.empty((C * N * 11,), dtype=torch.float32, device=self.device)
            o = 0
            views = []
            for w, shape in ((2, (C, N, 2)), (3, (C, N, 3)), (3, (C, N, 3)), (1, (C, N)), (1, (C, N)), (1, (C, N))):
                views.append(self._gflat[o:o + C * N * w].view(shape))
                o += C * N * w
            self.v_means2d, self.v_conics, self.v_colors, self.v_opacities, self.v_betas, self.v_depths = views
        # the depth gradients (last C N floats) are only produced by the 4- and 1-channel modes
        return self._gflat if self.channels != 3 else self._gflat[:self.C * self.N * 10]

    def grad_args(self, sl: Optional[slice] = None):
        """(v_rows, rows_form) of ubs_fused_project_bwd* for the frame composite_backward() last differentiated;
        sl = a row range (a view at an offset of a buffer owned by self, no temporary)."""
        rows = self.v_rows if sl is None else self.v_rows[:, sl]
        return ptr(rows), (1 if self.channels == 3 and self.grad_rows else 0)

    @torch.no_grad()
    def backward(self, records: Tensor, viewmats: Tensor, Ks: Tensor, cam_pos: Tensor, timestamps: Optional[Tensor],
                 backgrounds: Optional[Tensor], v_render_colors: Tensor, v_render_alphas: Tensor,
                 v_records: Optional[Tensor] = None, adam=None, opacity_reg: float = 0.0,
                 scale_reg: float = 0.0, activated: bool = False, query: Optional[Tensor] = None,
                 alphas: Optional[Tensor] = None, v_viewmats: Optional[Tensor] = None,
                 v_segments=None) -> Optional[Tensor]:
        """Gradient of the most recent forward() w.r.t. the packed records ([N, stride], same layout).
        Must be called before the next forward(): it reuses that frame's tile lists and screen-space records.
        With `adam` (a training.PackedAdam) the optimiser step is applied inside the projection-backward kernel:
        `records` and the moments are updated in place, no gradient buffer is produced and None is returned.
        v_viewmats: None, or a [C,4,4] tensor that receives the gradient of the world-to-camera matrices through the
        projection (what fully_fused_projection hands back with viewmats_requires_grad, _wrapper.py:898).
        v_segments: None, or seven tensors / Nones (mean [N,D], rgb [N,3], opacity [N], beta0 [N], beta_c [N,D-3],
        scale [N,D], l_triangle [N,M]): the gradient leaves the projection backward in the reference's separate layout
        instead of as records (the drop-in route; no v_records is produced, None is returned)."""
        lib, s = self.lib, torch.cuda.current_stream().cuda_stream
        C, N, D = self.C, self.N, self.D
        self.composite_backward(backgrounds, v_render_colors, v_render_alphas, alphas)
        if adam is not None:
            # single-GPU batch-1 training: projection backward + Adam in one launch, records updated in place
            import ctypes

            assert not activated and query is None, "the fused Adam update works on raw parameter records"
            adam.step_count += 1
            cols = (ctypes.c_double * adam.stride)(*adam.lr_columns())
            with self._stage("fused_project_bwd_adam"):
              check(lib.ubs_fused_project_bwd_adam(
                C, N, D, ptr(records), ptr(viewmats), ptr(Ks), ptr(cam_pos), ptr(timestamps), self.W, self.H,
                self.eps2d, 1 if self.aa else 0, ptr(self.radii), ptr(self.conics), *self.grad_args(),
                ptr(adam.exp_avg), ptr(adam.exp_avg_sq), ctypes.cast(cols, ctypes.c_void_p), adam.betas[0], adam.betas[1], adam.eps,
                adam.step_count, float(opacity_reg), float(scale_reg), ptr(self.status), s),
                "ubs_fused_project_bwd_adam")
            return None
        if v_segments is not None:
            assert len(v_segments) == 7 and all(t is None or (t.is_contiguous() and t.dtype == torch.float32 and
                                                               t.shape[0] == N) for t in v_segments)
            with self._stage("fused_project_bwd"):
                check(lib.ubs_fused_project_bwd_unpacked(
                    C, N, D, ptr(records), ptr(viewmats), ptr(Ks), ptr(cam_pos), ptr(timestamps), self.W, self.H,
                    self.eps2d, 1 if self.aa else 0, ptr(self.radii), ptr(self.conics), *self.grad_args(),
                    *[ptr(t) for t in v_segments], ptr(v_viewmats), 1 if activated else 0, ptr(query), ptr(self.status),
                    s), "ubs_fused_project_bwd_unpacked")
            return None
        if v_records is None:
            v_records = torch.empty_like(records)
        with self._stage("fused_project_bwd"):
            self.project_backward_rows(records, viewmats, Ks, cam_pos, timestamps, v_records, 0, N, activated, query,
                                       v_viewmats)
        return v_records

    @torch.no_grad()
    def composite_backward(self, backgrounds: Optional[Tensor], v_render_colors: Tensor, v_render_alphas: Tensor,
                           alphas: Optional[Tensor] = None):
        """First half of backward(): screen-space gradients of the most recent forward() into self.v_*.
        alphas: the frame's alpha image when forward() wrote it to `out` instead of self.render_alphas."""
        alphas = self.render_alphas if alphas is None else alphas
        assert alphas.shape == self.render_alphas.shape and alphas.is_contiguous()
        assert self.has_screen_space, "the frame was rendered with screen_space=False: nothing to differentiate"
        lib, s = self.lib, torch.cuda.current_stream().cuda_stream
        C, N = self.C, self.N
        ch = self.channels
        self._grad_buffers().zero_()  # one memset for all the screen-space gradient arrays
        v_rc, v_ra = v_render_colors.contiguous(), v_render_alphas.contiguous()  # locals: alive until after the launch
        assert v_rc.shape == (C, self.H, self.W, ch) and v_ra.shape == (C, self.H, self.W, 1)
        with self._stage("rasterize_bwd"):
          if ch == 3 and self.grad_rows:
            check(lib.ubs_rasterize_bwd_rows(
              C, N, ptr(self.n_isects), self.capacity, ptr(self.splats), ptr(backgrounds), None, self.W, self.H,
              self.tile_size, ptr(self.offsets), ptr(self.flatten_ids), ptr(alphas), ptr(self.last_ids), ptr(v_rc),
              ptr(v_ra), ptr(self.v_rows), ptr(self.status), s), "ubs_rasterize_bwd_rows")
          else:
            check(lib.ubs_rasterize_bwd_splats(
              C, N, ptr(self.n_isects), self.capacity, ptr(self.splats), None, ptr(backgrounds), None, ch, self.W,
              self.H, self.tile_size, ptr(self.offsets), ptr(self.flatten_ids), ptr(alphas), ptr(self.last_ids),
              ptr(v_rc), ptr(v_ra), ptr(self.v_means2d), ptr(self.v_conics), ptr(self.v_colors),
              ptr(self.v_opacities), ptr(self.v_betas), ptr(self.v_depths) if ch != 3 else None, s),
              "ubs_rasterize_bwd_splats")
            check(lib.ubs_pack_gradient_rows(
              C * N, ptr(self.v_means2d), ptr(self.v_depths) if ch != 3 else None, ptr(self.v_conics),
              ptr(self.v_opacities), ptr(self.v_betas), ptr(self.v_colors), ptr(self.v_rows), s),
              "ubs_pack_gradient_rows")

    @torch.no_grad()
    def project_backward_rows(self, records, viewmats, Ks, cam_pos, timestamps, v_records, begin: int, count: int,
                              activated: bool = False, query: Optional[Tensor] = None,
                              v_viewmats: Optional[Tensor] = None):
        """Second half of backward() for primitives [begin, begin + count): gradient records of those rows from
        self.v_*.  Row ranges are independent, so a caller can pipeline them against a collective (one camera)."""
        assert self.C == 1 or (begin == 0 and count == self.N), "row ranges need the [C, N] arrays to be [1, N]"
        assert v_viewmats is None or (begin == 0 and count == self.N and v_viewmats.shape == (self.C, 4, 4) and
                                      v_viewmats.is_contiguous() and v_viewmats.dtype == torch.float32)
        if count == 0:
            return
        sl = slice(begin, begin + count)
        # row slices of the [1, N, ...] arrays are views at an offset: no temporaries whose address could be recycled
        check(self.lib.ubs_fused_project_bwd(
            self.C, count, self.D, ptr(records[sl]), ptr(viewmats), ptr(Ks), ptr(cam_pos), ptr(timestamps), self.W,
            self.H, self.eps2d, 1 if self.aa else 0, ptr(self.radii[:, sl]), ptr(self.conics[:, sl]),
            *self.grad_args(sl), ptr(v_records[sl]), ptr(v_viewmats), 1 if activated else 0,
            ptr(None if query is None else query[sl]), ptr(self.status),
            torch.cuda.current_stream().cuda_stream), "ubs_fused_project_bwd")


class RenderQueue:
    """Throughput mode for streams of independent frames (camera-parallel rendering): `depth`
    rasterisers, each with its own CUDA stream, are used round-robin.  A frame is a chain of dependent kernels whose
    last and longest one (compositing) ends in partially filled waves while the first ones of the next frame are HBM /
    L2 bound; two frames in flight fill those gaps: 0.877 -> 0.813 ms per frame at 3M primitives / 1080p (three: 0.806).
    Latency per frame is unchanged.  Results of slot i are valid once its stream reaches the point of the call
    (`wait(slot)`, or `join()` to order the current stream behind everything queued)."""

    def __init__(self, make_rasterizer, depth: int = 2):
        self.rzs = [make_rasterizer() for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=rz.device) for rz in self.rzs]
        self.k = 0

    def render(self, records, viewmats, Ks, cam_pos, timestamps=None, backgrounds=None, **kw):
        """Queues one frame on the next slot; returns (slot, colors, alphas) -- buffers of that slot's rasteriser
        (or `out=`), overwritten when the slot comes round again."""
        slot = self.k % len(self.rzs)
        self.k += 1
        st = self.streams[slot]
        st.wait_stream(torch.cuda.current_stream())  # the caller's stream produced the inputs
        with torch.cuda.stream(st):
            rc, ra = self.rzs[slot].forward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, **kw)
        return slot, rc, ra

    def wait(self, slot: int):
        self.streams[slot].synchronize()

    def join(self):
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)


class HostPipeline:
    """Host-buffer front end of the renderer: camera parameters come from pinned host memory, the finished image
    goes back to pinned host memory, and `depth` frames are kept in flight so the device->host copy of frame k
    overlaps the kernels of frame k+1 (separate copy stream, double-buffered device and host images).  Given several
    rasterisers (`rz` a list), consecutive frames also alternate between them on their own streams (RenderQueue's
    overlap of one frame's compositing tail with the next frame's projection).

    Camera row layout (29 floats): viewmat 4x4 row-major | K 3x3 | camera centre xyz | timestamp."""

    CAM_FLOATS = 29

    def __init__(self, rz, depth: int = 2, copy_alpha: bool = False):
        # copy_alpha=False: the host gets what BetaModel.view / Scene.eval hand back -- the colour image
        # (scene/beta_model.py:827-831); the alpha plane stays on the device unless asked for
        self.rzs = list(rz) if isinstance(rz, (list, tuple)) else [rz]
        rz = self.rzs[0]
        self.render_streams = ([torch.cuda.Stream(device=rz.device) for _ in self.rzs] if len(self.rzs) > 1 else [None])
        self.rz, self.depth, self.copy_alpha = rz, depth, copy_alpha
        dev, C, H, W = rz.device, rz.C, rz.H, rz.W
        assert C == 1, "HostPipeline renders one camera per frame"
        self.cam_dev = [torch.empty((self.CAM_FLOATS,), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.img_dev = [(torch.empty((C, H, W, 3), dtype=torch.float32, device=dev),
                         torch.empty((C, H, W, 1), dtype=torch.float32, device=dev)) for _ in range(depth)]
        self.img_host = [(torch.empty((C, H, W, 3), dtype=torch.float32).pin_memory(),
                          torch.empty((C, H, W, 1), dtype=torch.float32).pin_memory()) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = [None] * depth  # event: D2H of slot finished
        self.k = 0

    def render_to_host(self, records: Tensor, cam_row_pinned: Tensor, backgrounds: Optional[Tensor] = None):
        """Queues one frame; returns the slot index whose host image will hold it after drain()/wait(slot)."""
        slot = self.k % self.depth
        rz = self.rzs[self.k % len(self.rzs)]
        main = self.render_streams[self.k % len(self.rzs)]
        self.k += 1
        if main is None:
            main = torch.cuda.current_stream()
        else:
            main.wait_stream(torch.cuda.current_stream())  # the caller's stream produced the records
        with torch.cuda.stream(main):
            if self.copied[slot] is not None:
                main.wait_event(self.copied[slot])  # the slot's device image is free again
            cam = self.cam_dev[slot]
            cam.copy_(cam_row_pinned, non_blocking=True)
            ts = cam[28:29] if rz.D == 7 else None
            rc, ra = rz.forward(records, cam[0:16].view(1, 4, 4), cam[16:25].view(1, 3, 3), cam[25:28].view(1, 3), ts,
                                backgrounds, out=self.img_dev[slot], screen_space=False)
            done = torch.cuda.Event()
            done.record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(done)
            self.img_host[slot][0].copy_(rc, non_blocking=True)
            if self.copy_alpha:
                self.img_host[slot][1].copy_(ra, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.copied[slot] = ev
        return slot

    def wait(self, slot: int):
        if self.copied[slot] is not None:
            self.copied[slot].synchronize()
        return self.img_host[slot]

    def drain(self):
        for ev in self.copied:
            if ev is not None:
                ev.synchronize()


class _FusedRender(torch.autograd.Function):
    """records -> (render_colors, render_alphas) with gradients to the packed records and the backgrounds."""

    @staticmethod
    def forward(ctx, records, rz, viewmats, Ks, cam_pos, timestamps, backgrounds, channels, activated, query):
        C, H, W = rz.C, rz.H, rz.W
        # fresh output tensors: autograd nodes downstream (the loss) save the image, and the rasteriser's own
        # buffers are overwritten by its next frame
        out = (torch.empty((C, H, W, channels), dtype=torch.float32, device=rz.device),
               torch.empty((C, H, W, 1), dtype=torch.float32, device=rz.device))
        rc, ra = rz.forward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, out=out, channels=channels,
                            activated=activated, query=query)
        ctx.rz, ctx.frame_id, ctx.activated = rz, rz.frame_id, activated
        ctx.save_for_backward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, query, ra)
        return rc, ra

    @staticmethod
    def backward(ctx, v_rc, v_ra):
        records, viewmats, Ks, cam_pos, timestamps, backgrounds, query, ra = ctx.saved_tensors
        rz = ctx.rz
        if rz.frame_id != ctx.frame_id:
            raise UbsError(
                "stale frame: this FusedRasterizer has rendered %d frame(s) since the one being differentiated; its tile "
                "lists and screen-space records are gone.  Call backward() before the next render on the same "
                "rasteriser, or give every frame in flight its own FusedRasterizer (ubs_b200.dropin keeps a pool)"
                % (rz.frame_id - ctx.frame_id))
        # the frame's pair count has normally reached the host by now: refuse to differentiate a truncated frame
        # (without blocking; if the copy is still in flight the device-side skip flag zeroes the gradient instead)
        if (rz._count_event is not None and rz._count_frame == ctx.frame_id and rz._count_event.query()
                and int(rz._host_status[0]) != 0):
            rz._poll_count()
            raise UbsError("the frame being differentiated was truncated (pair capacity exceeded); render it again -- "
                           "the buffers have been grown to %d pairs" % rz.capacity)
        v_view = torch.empty_like(viewmats) if ctx.needs_input_grad[2] else None
        v_records = rz.backward(records, viewmats, Ks, cam_pos, timestamps, backgrounds, v_rc, v_ra,
                                activated=ctx.activated, query=query, alphas=ra, v_viewmats=v_view)
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[6]:
            v_bg = (v_rc * (1.0 - ra)).sum(dim=(1, 2))
        return v_records, None, v_view, None, None, None, v_bg, None, None, None


def render(records: Tensor, rz: FusedRasterizer, viewmats: Tensor, Ks: Tensor, cam_pos: Tensor,
           timestamps: Optional[Tensor] = None, backgrounds: Optional[Tensor] = None, channels: int = 3,
           activated: bool = False, query: Optional[Tensor] = None):
    """Differentiable fused render: the equivalent of BetaModel.render (scene/beta_model.py:660-722) for C cameras
    at once.  Returns freshly allocated (render_colors [C,H,W,channels], render_alphas [C,H,W,1]).  backward() must
    run before `rz` renders its next frame (it raises otherwise)."""
    return _FusedRender.apply(records, rz, viewmats, Ks, cam_pos, timestamps, backgrounds, channels, activated, query)
