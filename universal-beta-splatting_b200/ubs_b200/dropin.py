"""Zero-edit drop-in route: the reference caller's own statements reach the fused kernels.

`BetaModel.render` / `.view` (scene/beta_model.py:660-722, 724-831) do

    get_rotation   -> l_triangle_to_rotmat(l_triangle[:, :3])                         (beta_model.py:123-125)
    get_covariance -> rot_scale_l_triangle_to_covar(rot, softplus(scale), l_triangle, rest_i, rest_j)   (:133-141)
    means, convs, opacities = cond_mean_convariance_opacity(mean, covar, opacity, beta[:, 1:], query)    (:154-159)
    rasterization(means[mask], ..., opacities.squeeze()[mask], betas, colors, viewmats, Ks, ..., covars=convs[mask])

Run eagerly that is K1, K2, K3, seven mask gathers, K5, the tile lists and K10 as separate launches with ~1 KB of
HBM traffic per primitive in between.  Here the three companion operators return *deferred* tensors instead
(`Deferred`, a storage-less torch.Tensor subclass that only remembers how it would be computed).  The few
operations the caller applies to them before `rasterization()` -- `[mask]`, `.squeeze()` -- stay deferred;
`rasterization()` then recognises its own conditioned tensors, packs the seven activated inputs of the chain into
records (ONE pass, csrc/pack.cu) and runs the fused kernels (`activated` records, the caller's own `query`), with one
autograd node whose backward hands gradients to exactly the tensors the chain received (the projection backward writes
the seven gradient tensors itself, ubs_fused_project_bwd_unpacked).  Any other use of a deferred
tensor (arithmetic, printing, indexing by anything but a primitive mask, ...) materialises it through the stand-alone
operators -- same values as before, just not fused -- so nothing the reference could do stops working.

`UBS_DROPIN_FUSED=0` in the environment turns the deferral off (every operator call is eager).
"""
import os
import threading
from typing import Dict, List, Optional

import torch
from torch import Tensor

from . import _lib, fused, ops
from ._lib import check, ptr

ENABLED = os.environ.get("UBS_DROPIN_FUSED", "1") != "0"
MAX_FRAMES_IN_FLIGHT = 8  # rendered-but-not-yet-differentiated frames per (scene size, image size): train.py's batch loop

_STATS = {"fused": 0, "fallback": 0, "materialized": 0}


def stats() -> Dict[str, int]:
    """Counters of this process: rasterization() calls that took the fused route / fell back, and deferred tensors
    that had to be materialised."""
    return dict(_STATS)


# ----------------------------------------------------------------------------------------------------------------
# deferred tensors
# ----------------------------------------------------------------------------------------------------------------
class _Node:
    """One deferred operator call: its real inputs and, once somebody needs them, its real outputs."""

    def __init__(self, kind: str, inputs: tuple, compute):
        self.kind, self.inputs, self._compute, self._real = kind, inputs, compute, None

    def real(self):
        if self._real is None:
            _STATS["materialized"] += 1
            self._real = self._compute()
        return self._real


# metadata getters that never need the values
_META_NAMES = ("shape", "device", "dtype", "requires_grad", "ndim", "is_cuda", "layout", "is_sparse", "is_quantized",
               "is_meta", "names", "grad_fn", "is_leaf", "grad")
_META_FUNCS = set()
for _n in _META_NAMES:
    _d = getattr(torch.Tensor, _n, None)
    if _d is not None and hasattr(_d, "__get__"):
        _META_FUNCS.add(_d.__get__)
for _n in ("size", "dim", "numel", "__len__", "is_contiguous", "is_floating_point", "get_device", "nelement", "ndimension",
           "element_size", "is_complex", "data_ptr"):
    _META_FUNCS.add(getattr(torch.Tensor, _n))
_META_FUNCS.discard(torch.Tensor.data_ptr)  # an address needs real storage


class Deferred(torch.Tensor):
    """A tensor that has not been computed yet: output `index` of `node`, followed by `views` (a tuple of
    ("mask", bool_tensor) / ("squeeze",) steps)."""

    @staticmethod
    def __new__(cls, node: _Node, index: int, shape, device, requires_grad: bool, views=()):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=torch.float32, device=device,
                                                requires_grad=False)
        t._node, t._index, t._views, t._rg = node, index, tuple(views), bool(requires_grad)
        return t

    def materialize(self) -> Tensor:
        out = self._node.real()
        out = out[self._index] if isinstance(out, tuple) else out
        for v in self._views:
            out = out[v[1]] if v[0] == "mask" else out.squeeze()
        return out

    def __repr__(self):
        return "Deferred(%s[%d], shape=%s)" % (self._node.kind, self._index, tuple(self.shape))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _META_FUNCS:
            if func == torch.Tensor.requires_grad.__get__:
                return args[0]._rg
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        self = args[0] if args else None
        if isinstance(self, Deferred) and not kwargs:
            # the two things the reference caller does to the conditioned tensors before rasterization()
            if func is torch.Tensor.__getitem__ and len(args) == 2 and _is_prim_mask(args[1], self):
                mask = args[1]
                with torch._C.DisableTorchFunctionSubclass():
                    shape = tuple(self.shape)
                # The kept count is only known on the device, so a masked deferred tensor reports the UNMASKED leading
                # size (an upper bound) until it is materialised; rasterization() takes the true count from the
                # caller's own colours / betas, which were gathered eagerly with the same mask.
                return Deferred(self._node, self._index, shape, self.device, self._rg, self._views + (("mask", mask),))
            if func in (torch.Tensor.squeeze, torch.squeeze) and len(args) == 1:
                with torch._C.DisableTorchFunctionSubclass():
                    shape = tuple(d for d in self.shape if d != 1)
                return Deferred(self._node, self._index, shape, self.device, self._rg, self._views + (("squeeze",),))
        # anything else: compute the real tensors and carry on
        real_args = _tree_materialize(args)
        real_kwargs = _tree_materialize(kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*real_args, **real_kwargs)


    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # Only reached when an ATen operator sees a deferred tensor without passing __torch_function__ (C++ callers):
        # same rule -- compute the real tensors and carry on.
        return func(*_tree_materialize(args), **_tree_materialize(kwargs or {}))


def _is_prim_mask(index, t: "Deferred") -> bool:
    return (isinstance(index, Tensor) and not isinstance(index, Deferred) and index.dtype == torch.bool
            and index.dim() == 1 and index.shape[0] == t.shape[0])


def _tree_materialize(x):
    if isinstance(x, Deferred):
        return x.materialize()
    if isinstance(x, (list, tuple)):
        return type(x)(_tree_materialize(v) for v in x)
    if isinstance(x, dict):
        return {k: _tree_materialize(v) for k, v in x.items()}
    return x


def materialize(x):
    """Real tensor(s) for deferred ones (identity on everything else)."""
    return _tree_materialize(x)


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, Tensor) and (t._rg if isinstance(t, Deferred) else t.requires_grad)
                                           for t in ts)


# ----------------------------------------------------------------------------------------------------------------
# the three companion operators, deferred (scene/beta_model.py:18-22 imports them from gsplat.cuda._wrapper)
# ----------------------------------------------------------------------------------------------------------------
def _deferrable(*ts) -> bool:
    return ENABLED and all(isinstance(t, Tensor) and not isinstance(t, Deferred) and t.is_cuda and t.dtype == torch.float32
                           for t in ts)


def l_triangle_to_rotmat(l_triangle: Tensor) -> Tensor:
    """cuda/_wrapper.py:34-36.  Deferred: [N,3] -> [N,3,3]."""
    if not _deferrable(l_triangle) or l_triangle.dim() != 2 or l_triangle.shape[1] != 3:
        return ops.l_triangle_to_rotmat(materialize(l_triangle))
    node = _Node("rotmat", (l_triangle,), lambda: ops.l_triangle_to_rotmat(l_triangle))
    return Deferred(node, 0, (l_triangle.shape[0], 3, 3), l_triangle.device, _needs_grad(l_triangle))


def rot_scale_l_triangle_to_covar(rot: Tensor, scale: Tensor, l_triangle: Tensor, rest_i: Tensor, rest_j: Tensor,
                                  spatial_block: bool = False) -> Tensor:
    """cuda/_wrapper.py:39-55.  Deferred when `rot` is the deferred rotation of this very `l_triangle`'s first three
    columns (get_covariance, scene/beta_model.py:133-141) and the full D x D matrix is asked for."""
    ok = (isinstance(rot, Deferred) and rot._node.kind == "rotmat" and not rot._views and not spatial_block
          and _deferrable(scale, l_triangle) and scale.dim() == 2 and scale.shape[1] in (6, 7)
          and l_triangle.dim() == 2 and l_triangle.shape == (scale.shape[0], scale.shape[1] * (scale.shape[1] - 1) // 2))
    if ok:
        src = rot._node.inputs[0]
        ok = (src.shape == (l_triangle.shape[0], 3) and src.data_ptr() == l_triangle.data_ptr()
              and src.stride() == (l_triangle.stride(0), l_triangle.stride(1)) and src._version == l_triangle._version)
    if not ok:
        return ops.rot_scale_l_triangle_to_covar(materialize(rot), scale, l_triangle, rest_i, rest_j, spatial_block)
    D = scale.shape[1]
    node = _Node("covar", (scale, l_triangle, rot),
                 lambda: ops.rot_scale_l_triangle_to_covar(rot.materialize(), scale, l_triangle, rest_i, rest_j, False))
    node.rest = (rest_i, rest_j)
    return Deferred(node, 0, (scale.shape[0], D, D), scale.device, _needs_grad(scale, l_triangle))


def cond_mean_convariance_opacity(means: Tensor, covars: Tensor, opacities: Tensor, betas: Tensor, query: Tensor):
    """cuda/_wrapper.py:18-31.  Deferred when `covars` is a deferred covariance."""
    ok = (isinstance(covars, Deferred) and covars._node.kind == "covar" and not covars._views
          and _deferrable(means, opacities, betas, query) and means.dim() == 2
          and means.shape == (covars.shape[0], covars.shape[1]) and opacities.shape == (means.shape[0], 1)
          and betas.shape == (means.shape[0], means.shape[1] - 3) and query.shape == betas.shape)
    if not ok:
        return ops.cond_mean_convariance_opacity(materialize(means), materialize(covars), materialize(opacities),
                                                 materialize(betas), materialize(query))
    N = means.shape[0]
    node = _Node("cond", (means, covars, opacities, betas, query),
                 lambda: ops.cond_mean_convariance_opacity(means, covars.materialize(), opacities, betas, query))
    rg = _needs_grad(means, covars, opacities, betas)
    dev = means.device
    return (Deferred(node, 0, (N, 3), dev, rg), Deferred(node, 1, (N, 3, 3), dev, rg), Deferred(node, 2, (N, 1), dev, rg))


# ----------------------------------------------------------------------------------------------------------------
# frames in flight: one rasteriser (tile lists, screen-space records, packed records) per rendered frame that still
# awaits its backward -- train.py:111-128 renders `batch_size` views before one backward()
# ----------------------------------------------------------------------------------------------------------------
class _Slot:
    def __init__(self, rz: fused.FusedRasterizer):
        self.rz, self.busy = rz, False
        self.records = torch.empty((rz.N, fused.record_stride(rz.D)), dtype=torch.float32, device=rz.device)
        self.v_records = None


class _Lease:
    """Marks a slot busy until the autograd node that owns it is done (backward ran) or dies (graph freed)."""

    def __init__(self, slot: _Slot):
        self.slot = slot
        slot.busy = True

    def release(self):
        if self.slot is not None:
            self.slot.busy = False
            self.slot = None

    def __del__(self):
        self.release()


_POOL: Dict[tuple, List[_Slot]] = {}
_POOL_LOCK = threading.Lock()  # the viewer thread renders while the training thread does (train.py:95-98,177)


def _acquire(key, make) -> Optional[_Slot]:
    with _POOL_LOCK:
        if key not in _POOL and len(_POOL) >= 6:  # a viewer resizing its window: keep the cache bounded
            for k in [k for k, v in _POOL.items() if not any(s.busy for s in v)][:1]:
                del _POOL[k]
        slots = _POOL.setdefault(key, [])
        for s in slots:
            if not s.busy:
                return s
        if len(slots) >= MAX_FRAMES_IN_FLIGHT:
            return None
        s = _Slot(make())
        slots.append(s)
        return s


def clear_pool():
    with _POOL_LOCK:
        _POOL.clear()


class _DropinRender(torch.autograd.Function):
    """(mean, rgb, opacity, beta0, beta_c, scale, l_triangle, backgrounds) -> image, through activated records."""

    @staticmethod
    def forward(ctx, mean, rgb, opacity, beta0, beta_c, scale, ltri, backgrounds, query, slot, viewmats, Ks, channels,
                prim_mask, track):
        lib, rz = _lib.load(), slot.rz
        N, D = rz.N, rz.D
        srcs = [t.contiguous() for t in (mean, rgb, opacity, beta0, beta_c, scale, ltri)]  # alive until after the launch
        check(lib.ubs_pack_records(N, D, *[ptr(t) for t in srcs], ptr(slot.records),
                                   torch.cuda.current_stream().cuda_stream), "ubs_pack_records")
        out = (torch.empty((rz.C, rz.H, rz.W, channels), dtype=torch.float32, device=rz.device),
               torch.empty((rz.C, rz.H, rz.W, 1), dtype=torch.float32, device=rz.device))
        rc, ra = rz.forward(slot.records, viewmats, Ks, None, None, backgrounds, prim_mask=prim_mask, out=out,
                            channels=channels, activated=True, query=query)
        if track:
            ctx.lease = _Lease(slot)
            ctx.slot, ctx.frame_id = slot, rz.frame_id
            ctx.save_for_backward(viewmats, Ks, backgrounds, query, ra)
        return rc, ra

    @staticmethod
    def backward(ctx, v_rc, v_ra):
        viewmats, Ks, backgrounds, query, ra = ctx.saved_tensors
        slot = ctx.slot
        rz = slot.rz
        if rz.frame_id != ctx.frame_id:
            raise _lib.UbsError("stale frame in the drop-in route (rasteriser slot reused before its backward)")
        N, D = rz.N, rz.D
        need = ctx.needs_input_grad
        dev = rz.device
        f32 = torch.float32
        M = D * (D - 1) // 2
        g = [torch.empty((N, D), dtype=f32, device=dev) if need[0] else None,
             torch.empty((N, 3), dtype=f32, device=dev) if need[1] else None,
             torch.empty((N,), dtype=f32, device=dev) if need[2] else None,
             torch.empty((N,), dtype=f32, device=dev) if need[3] else None,
             torch.empty((N, D - 3), dtype=f32, device=dev) if need[4] else None,
             torch.empty((N, D), dtype=f32, device=dev) if need[5] else None,
             torch.empty((N, M), dtype=f32, device=dev) if need[6] else None]
        # the projection backward writes the seven gradients itself (no gradient record buffer, no unpack pass)
        rz.backward(slot.records, viewmats, Ks, None, None, backgrounds, v_rc, v_ra, activated=True, query=query,
                    alphas=ra, v_segments=g)
        v_bg = None
        if backgrounds is not None and need[7]:
            v_bg = (v_rc * (1.0 - ra)).sum(dim=(1, 2))
        ctx.lease.release()
        return (g[0], g[1], g[2], g[3], g[4], g[5], g[6], v_bg, None, None, None, None, None, None, None)


class LazyMeta(dict):
    """The `meta` dict of rasterization() (submodules/gsplat/rendering.py:107-120,161-174).  The exactly-sized pair
    arrays need the pair count on the host -- the one synchronisation of the reference's isect_tiles
    (isect_tiles.cu:180-181) -- so they are produced on first access only; the caller (scene/beta_model.py:716-722,823)
    reads `means2d` and `radii`."""

    _LAZY = ("isect_ids", "flatten_ids")

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._thunks = {}

    def defer(self, key, thunk):
        self._thunks[key] = thunk
        super().__setitem__(key, None)

    def __getitem__(self, key):
        if key in self._thunks:
            super().__setitem__(key, self._thunks.pop(key)())
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        for k in list(self._thunks):
            self[k]
        return super().items()

    def values(self):
        for k in list(self._thunks):
            self[k]
        return super().values()


def try_fused_rasterization(means, opacities, betas, colors, viewmats, Ks, width, height, near_plane, far_plane,
                            radius_clip, eps2d, tile_size, backgrounds, render_mode, rasterize_mode, covars):
    """The fused route of rasterization(), or None when its preconditions do not hold (the caller then materialises
    the deferred tensors and runs the operator chain).  Preconditions: means / opacities / covars are the three
    deferred outputs of ONE cond_mean_convariance_opacity call, all three behind the same (or no) primitive mask,
    opacities squeezed; colours [n,3] and betas [n] real tensors; one of the six UBS render modes; tile size 16."""
    if not (isinstance(means, Deferred) and isinstance(opacities, Deferred) and isinstance(covars, Deferred)):
        return None
    node = means._node
    if not (node.kind == "cond" and opacities._node is node and covars._node is node
            and (means._index, covars._index, opacities._index) == (0, 1, 2)):
        return None

    def mask_of(t, allow_squeeze):
        mask, squeezed = None, False
        for v in t._views:
            if v[0] == "mask" and mask is None:
                mask = v[1]
            elif v[0] == "squeeze" and allow_squeeze and not squeezed and mask is None:
                squeezed = True
            else:
                return False, None
        return True, mask

    ok_m, mask = mask_of(means, False)
    ok_c, mask_c = mask_of(covars, False)
    ok_o, mask_o = mask_of(opacities, True)
    if not (ok_m and ok_c and ok_o) or mask_c is not mask or mask_o is not mask:
        return None
    if not any(v[0] == "squeeze" for v in opacities._views):
        return None
    mean, covar_d, opac, beta_c, query = node.inputs
    scale, ltri, _ = covar_d._node.inputs
    N, D = mean.shape
    C = viewmats.shape[0]
    channels = {"RGB": 3, "RGB+D": 4, "RGB+ED": 4, "Depth": 1, "EDepth": 1, "Normal": 1}.get(render_mode)
    if channels is None or tile_size != 16 or rasterize_mode not in ("classic", "antialiased"):
        return None
    if not (isinstance(colors, Tensor) and not isinstance(colors, Deferred) and colors.dim() == 2 and colors.shape[1] == 3
            and isinstance(betas, Tensor) and not isinstance(betas, Deferred) and betas.dim() == 1
            and colors.shape[0] == betas.shape[0] and colors.is_cuda and colors.dtype == torch.float32
            and betas.dtype == torch.float32):
        return None
    if viewmats.requires_grad and torch.is_grad_enabled():
        return None  # the fused backward produces no camera gradient
    n_kept = colors.shape[0]
    if mask is None and n_kept != N:
        return None
    grad = _needs_grad(mean, opac, beta_c, scale, ltri, colors, betas, backgrounds)
    partial = mask is not None and n_kept != N
    if partial and grad:
        return None  # a filtered, differentiated render (never done by the reference): operator chain
    dev = mean.device
    if backgrounds is not None and tuple(backgrounds.shape) != (C, 3):
        return None
    if partial:
        # viewer filter (scene/beta_model.py:729-755, 797-802): the kernel skips masked-out primitives itself; the
        # caller's compacted colours / betas go back to their rows
        full_c = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        full_b = torch.zeros((N,), dtype=torch.float32, device=dev)
        full_c[mask] = colors
        full_b[mask] = betas
        colors, betas = full_c, full_b

    aa = rasterize_mode == "antialiased"
    key = (D, N, int(width), int(height), C, float(near_plane), float(far_plane), float(radius_clip), float(eps2d), aa,
           str(dev))
    slot = _acquire(key, lambda: fused.FusedRasterizer(D, N, int(width), int(height), n_cams=C, device=dev, eps2d=eps2d,
                                                       near_plane=near_plane, far_plane=far_plane,
                                                       radius_clip=radius_clip, antialiased=aa))
    if slot is None:
        return None
    bg = None
    if backgrounds is not None:
        # depth channel / depth-only modes composite over a zero background (rendering.py:131-142)
        if channels == 4:
            bg = torch.cat([backgrounds, torch.zeros(C, 1, device=dev)], dim=-1)
        elif channels == 1:
            bg = torch.zeros(C, 1, device=dev)
        else:
            bg = backgrounds
        bg = bg.contiguous()
    vm, K = viewmats.contiguous(), Ks.contiguous()
    rc, ra = _DropinRender.apply(mean, colors, opac.reshape(N), betas, beta_c, scale, ltri, bg, query.contiguous(), slot,
                                 vm, K, channels, mask if partial else None, grad)
    rz = slot.rz
    _STATS["fused"] += 1

    def sel(t):  # per-primitive outputs of a filtered render are returned for the kept primitives, like the reference
        return t[:, mask] if partial else t

    meta = LazyMeta({
        "camera_ids": None, "primitive_ids": None, "radii": sel(rz.radii), "means2d": sel(rz.means2d),
        "depths": sel(rz.depths), "conics": sel(rz.conics), "opacities": sel(rz.opacities), "betas": sel(rz.betas),
        "tile_width": rz.tw, "tile_height": rz.th, "tiles_per_gauss": sel(rz.tiles_per_gauss),
        "isect_offsets": rz.offsets, "width": width, "height": height, "tile_size": tile_size, "n_cameras": C})
    if partial:
        # flatten ids of the reference index the compacted primitive list
        remap = torch.cumsum(mask.to(torch.int64), 0) - 1

        def flat():
            n = rz.last_pair_count()
            f = rz.flatten_ids[:n].to(torch.int64)
            return ((f // N) * n_kept + remap[f % N]).to(torch.int32)
    else:
        def flat():
            return rz.flatten_ids[:rz.last_pair_count()]
    meta.defer("isect_ids", lambda: rz.isect_ids[:rz.last_pair_count()])
    meta.defer("flatten_ids", flat)
    return rc, ra, meta
