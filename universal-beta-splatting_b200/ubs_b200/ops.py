"""Torch-level operators over the C-ABI (device memory and streams come from PyTorch; the work is in the library).

Names, arguments and return values mirror the reference's operator API
(submodules/gsplat/cuda/_wrapper.py:18-55,185-491) so that parity tests read like calls into the reference.
Every function launches on torch's current CUDA stream and raises if the inputs are not CUDA float32 tensors:
there is no CPU path.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr

SUPPORTED_CHANNELS = (1, 2, 3, 4, 8, 16)  # one compositing launch; wider colour vectors are chunked


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _storage_types(fn):
    """The reference instantiates K1-K4 for half and bfloat16 as well (AT_DISPATCH_FLOATING_TYPES_AND2,
    cond_mean_convariance_opacity_fwd.cu:330, rot_scale_l_triangle_to_covar_fwd.cu:216, l_triagnle_to_rotmat_fwd.cu:50):
    storage in the reduced type, transcendentals in float.  Here such inputs are widened once, the float32 kernels run,
    and outputs (and, through autograd's cast nodes, gradients) come back in the input's type -- every intermediate is
    at least as precise as the reference's.  float64 has no instantiation (the caller is float32): rejected by _req."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dt = next((a.dtype for a in args if isinstance(a, Tensor) and a.is_floating_point()), None)
        if dt not in (torch.float16, torch.bfloat16):
            return fn(*args, **kw)
        out = fn(*[a.float() if isinstance(a, Tensor) and a.dtype == dt else a for a in args], **kw)
        return tuple(o.to(dt) for o in out) if isinstance(out, tuple) else out.to(dt)

    return wrapped


def _req(t: Tensor, name: str, dtype=torch.float32) -> Tensor:
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (ubs_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


# --------------------------------------------------------------------------------------------------------------
# K5/K6 projection
# --------------------------------------------------------------------------------------------------------------
def projection_fwd(means, covars, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                   calc_compensations):
    lib = _lib.load()
    means, covars = _req(means, "means"), _req(covars, "covars")
    viewmats, Ks = _req(viewmats, "viewmats"), _req(Ks, "Ks")
    C, N = viewmats.shape[0], means.shape[0]
    dev = means.device
    radii = torch.empty((C, N), dtype=torch.int32, device=dev)
    means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((C, N), dtype=torch.float32, device=dev)
    conics = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
    comps = torch.empty((C, N), dtype=torch.float32, device=dev) if calc_compensations else None
    check(lib.ubs_projection_fwd(C, N, ptr(means), ptr(covars), ptr(viewmats), ptr(Ks), int(width), int(height),
                                 float(eps2d), float(near_plane), float(far_plane), float(radius_clip), ptr(radii),
                                 ptr(means2d), ptr(depths), ptr(conics), ptr(comps), _stream()),
          "ubs_projection_fwd")
    return radii, means2d, depths, conics, comps


def projection_bwd(means, covars, viewmats, Ks, width, height, eps2d, radii, conics, compensations, v_means2d,
                   v_depths, v_conics, v_compensations, viewmats_requires_grad):
    lib = _lib.load()
    C, N = viewmats.shape[0], means.shape[0]
    dev = means.device
    v_means = torch.empty((N, 3), dtype=torch.float32, device=dev)
    v_covars = torch.empty((N, 6), dtype=torch.float32, device=dev)
    v_viewmats = torch.empty((C, 4, 4), dtype=torch.float32, device=dev) if viewmats_requires_grad else None
    # contiguous copies are bound to locals so that they outlive the launch (upstream gradients are often views)
    v_means2d, v_depths, v_conics = _req(v_means2d, "v_means2d"), _req(v_depths, "v_depths"), _req(v_conics, "v_conics")
    v_comp = None if v_compensations is None else _req(v_compensations, "v_compensations")
    check(lib.ubs_projection_bwd(C, N, ptr(means), ptr(covars), ptr(viewmats), ptr(Ks), int(width), int(height),
                                 float(eps2d), ptr(radii), ptr(conics), ptr(compensations),
                                 ptr(v_means2d), ptr(v_depths), ptr(v_conics), ptr(v_comp),
                                 ptr(v_means), ptr(v_covars), ptr(v_viewmats), _stream()),
          "ubs_projection_bwd")
    return v_means, v_covars, v_viewmats


class _FullyFusedProjection(torch.autograd.Function):
    """Mirror of the reference's _FullyFusedProjection (cuda/_wrapper.py:807-925), covars path."""

    @staticmethod
    def forward(ctx, means, covars, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                calc_compensations):
        radii, means2d, depths, conics, comps = projection_fwd(
            means, covars, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip, calc_compensations)
        ctx.save_for_backward(means, covars, viewmats, Ks, radii, conics, comps)
        ctx.dims = (width, height, eps2d)
        ctx.mark_non_differentiable(radii)
        return radii, means2d, depths, conics, comps

    @staticmethod
    def backward(ctx, v_radii, v_means2d, v_depths, v_conics, v_compensations):
        means, covars, viewmats, Ks, radii, conics, comps = ctx.saved_tensors
        width, height, eps2d = ctx.dims
        v_means, v_covars, v_viewmats = projection_bwd(
            means.contiguous(), covars.contiguous(), viewmats.contiguous(), Ks.contiguous(), width, height, eps2d,
            radii, conics, comps, v_means2d, v_depths, v_conics, v_compensations if comps is not None else None,
            ctx.needs_input_grad[2])
        return (v_means if ctx.needs_input_grad[0] else None, v_covars if ctx.needs_input_grad[1] else None,
                v_viewmats, None, None, None, None, None, None, None, None)


def fully_fused_projection(
    means: Tensor,  # [N, 3]
    covars: Optional[Tensor],  # [N, 6]
    quats: Optional[Tensor],
    scales: Optional[Tensor],
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    sparse_grad: bool = False,
    calc_compensations: bool = False,
    ortho: bool = False,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Optional[Tensor]]:
    """Same call as the reference (cuda/_wrapper.py:185-282).  Universal Beta Splatting always passes `covars`
    (rendering.py:48-56); the quaternion/ortho branches of upstream gsplat are dead code on this path and are
    rejected explicitly."""
    C, N = viewmats.size(0), means.size(0)
    assert means.size() == (N, 3), means.size()
    assert viewmats.size() == (C, 4, 4), viewmats.size()
    assert Ks.size() == (C, 3, 3), Ks.size()
    if covars is None:
        raise NotImplementedError("ubs_b200 projects from covariances only (the quats/scales path is unused by UBS)")
    if ortho:
        raise NotImplementedError("orthographic projection is not on the UBS hot path")
    assert covars.size() == (N, 6), covars.size()
    return _FullyFusedProjection.apply(_req(means, "means"), _req(covars, "covars"), _req(viewmats, "viewmats"),
                                       _req(Ks, "Ks"), width, height, eps2d, near_plane, far_plane, radius_clip,
                                       calc_compensations)


# --------------------------------------------------------------------------------------------------------------
# K7-K9 tile intersection
# --------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def isect_tiles(
    means2d: Tensor,  # [C, N, 2]
    radii: Tensor,  # [C, N]
    depths: Tensor,  # [C, N]
    tile_size: int,
    tile_width: int,
    tile_height: int,
    sort: bool = True,
    n_cameras: Optional[int] = None,
    camera_ids: Optional[Tensor] = None,
    primitive_ids: Optional[Tensor] = None,
    return_offsets: bool = False,
    method: str = "onesweep",
):
    """Same call and results as the reference (cuda/_wrapper.py:285-343): exactly-sized, sorted isect_ids /
    flatten_ids, which costs one device->host read of the pair count (the reference has the same sync,
    isect_tiles.cu:180-181).  With return_offsets=True the tile offsets are produced in the same pass.
    method="onesweep": emit + global stable radix sort (the reference's structure); method="bin": tile binning +
    per-tile segment sort (csrc/bin_sort.cu; sorted output only, depths must be >= 0) -- identical results."""
    assert camera_ids is None and primitive_ids is None, "packed mode is not used by UBS"
    assert method in ("onesweep", "bin")
    lib = _lib.load()
    C, N, _ = means2d.shape
    assert means2d.shape == (C, N, 2), means2d.size()
    assert radii.shape == (C, N), radii.size()
    assert depths.shape == (C, N), depths.size()
    means2d, depths = _req(means2d, "means2d"), _req(depths, "depths")
    radii = _req(radii, "radii", torch.int32)
    dev = means2d.device
    tiles_per_gauss = torch.empty((C, N), dtype=torch.int32, device=dev)
    n_isects_dev = torch.empty((1,), dtype=torch.int64, device=dev)
    if method == "bin":
        assert sort, "the binned route only produces the sorted list"
        offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=dev)
        isect_ids = flatten_ids = None
        # pass 1 (capacity 0) counts; pass 2 fills exactly-sized arrays
        for cap in (0, None):
            if cap is None:
                cap = int(n_isects_dev.item())
                isect_ids = torch.empty((cap,), dtype=torch.int64, device=dev)
                flatten_ids = torch.empty((cap,), dtype=torch.int32, device=dev)
                if cap == 0:
                    break
            ws = torch.empty((lib.ubs_isect_bin_workspace_bytes(C, tile_width, tile_height, cap),), dtype=torch.uint8,
                             device=dev)
            check(lib.ubs_isect_bin_sort(C, N, ptr(means2d), ptr(radii), ptr(depths), tile_size, tile_width,
                                         tile_height, 0, ptr(tiles_per_gauss), cap, ptr(n_isects_dev), ptr(isect_ids),
                                         ptr(flatten_ids), ptr(offsets), None, ptr(ws), ws.numel(), _stream()),
                  "ubs_isect_bin_sort")
        if return_offsets:
            return tiles_per_gauss, isect_ids, flatten_ids, offsets
        return tiles_per_gauss, isect_ids, flatten_ids
    ws0 = torch.empty((lib.ubs_isect_workspace_bytes(C * N, 0),), dtype=torch.uint8, device=dev)
    check(lib.ubs_isect_count(C, N, ptr(means2d), ptr(radii), tile_size, tile_width, tile_height,
                              ptr(tiles_per_gauss), ptr(n_isects_dev), ptr(ws0), ws0.numel(), _stream()),
          "ubs_isect_count")
    n_isects = int(n_isects_dev.item())  # the one host sync of the compat path
    isect_ids = torch.empty((n_isects,), dtype=torch.int64, device=dev)
    flatten_ids = torch.empty((n_isects,), dtype=torch.int32, device=dev)
    offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=dev) if return_offsets else None
    ws = torch.empty((lib.ubs_isect_workspace_bytes(C * N, n_isects),), dtype=torch.uint8, device=dev)
    ws[: ws0.numel()].copy_(ws0)  # scanned block offsets from the count phase
    check(lib.ubs_isect_emit_sort(C, N, ptr(means2d), ptr(radii), ptr(depths), tile_size, tile_width, tile_height,
                                  1 if sort else 0, ptr(tiles_per_gauss), ptr(n_isects_dev), n_isects,
                                  ptr(isect_ids), ptr(flatten_ids), ptr(offsets), None, ptr(ws), ws.numel(),
                                  _stream()),
          "ubs_isect_emit_sort")
    if return_offsets:
        return tiles_per_gauss, isect_ids, flatten_ids, offsets
    return tiles_per_gauss, isect_ids, flatten_ids


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """Same call as the reference (cuda/_wrapper.py:346-363)."""
    lib = _lib.load()
    isect_ids = _req(isect_ids, "isect_ids", torch.int64)
    offsets = torch.empty((n_cameras, tile_height, tile_width), dtype=torch.int32, device=isect_ids.device)
    check(lib.ubs_isect_offset_encode(isect_ids.numel(), ptr(isect_ids), n_cameras, tile_width, tile_height,
                                      ptr(offsets), _stream()),
          "ubs_isect_offset_encode")
    return offsets


@torch.no_grad()
def radix_sort_pairs(keys: Tensor, vals: Tensor, begin_bit: int = 0, end_bit: int = 64):
    """Stable ascending sort of (int64 key, int32 value) pairs on key bits [begin_bit, end_bit)."""
    lib = _lib.load()
    keys, vals = _req(keys, "keys", torch.int64).clone(), _req(vals, "vals", torch.int32).clone()
    n = keys.numel()
    n_dev = torch.tensor([n], dtype=torch.int64, device=keys.device)
    keys_out, vals_out = torch.empty_like(keys), torch.empty_like(vals)
    ws = torch.empty((lib.ubs_radix_sort_workspace_bytes(n),), dtype=torch.uint8, device=keys.device)
    check(lib.ubs_radix_sort_pairs(ptr(n_dev), n, ptr(keys), ptr(vals), ptr(keys_out), ptr(vals_out), begin_bit,
                                   end_bit, ptr(ws), ws.numel(), _stream()),
          "ubs_radix_sort_pairs")
    return keys_out, vals_out


# --------------------------------------------------------------------------------------------------------------
# K10/K11 compositing
# --------------------------------------------------------------------------------------------------------------
def rasterize_fwd(means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size,
                  isect_offsets, flatten_ids, n_isects_dev=None):
    lib = _lib.load()
    C, N = isect_offsets.shape[0], means2d.shape[1]
    dev = means2d.device
    ch = colors.shape[-1]
    if n_isects_dev is None:
        n_isects_dev = torch.tensor([flatten_ids.numel()], dtype=torch.int64, device=dev)
    render_colors = torch.empty((C, height, width, ch), dtype=torch.float32, device=dev)
    render_alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
    last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
    check(lib.ubs_rasterize_fwd(C, N, ptr(n_isects_dev), flatten_ids.numel(), ptr(means2d), ptr(conics), ptr(colors), ptr(opacities),
                                ptr(betas), ptr(backgrounds), ptr(masks), ch, int(width), int(height), int(tile_size),
                                ptr(isect_offsets), ptr(flatten_ids), ptr(render_colors), ptr(render_alphas),
                                ptr(last_ids), _stream()),
          "ubs_rasterize_fwd")
    return render_colors, render_alphas, last_ids


def rasterize_bwd(means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size,
                  isect_offsets, flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas,
                  n_isects_dev=None):
    lib = _lib.load()
    C, N = isect_offsets.shape[0], means2d.shape[1]
    dev = means2d.device
    ch = colors.shape[-1]
    if n_isects_dev is None:
        n_isects_dev = torch.tensor([flatten_ids.numel()], dtype=torch.int64, device=dev)
    v_means2d = torch.zeros_like(means2d)
    v_conics = torch.zeros_like(conics)
    v_colors = torch.zeros_like(colors)
    v_opacities = torch.zeros_like(opacities)
    v_betas = torch.zeros_like(betas)
    v_rc, v_ra = _req(v_render_colors, "v_render_colors"), _req(v_render_alphas, "v_render_alphas")  # outlive the launch
    check(lib.ubs_rasterize_bwd(C, N, ptr(n_isects_dev), flatten_ids.numel(), ptr(means2d), ptr(conics), ptr(colors), ptr(opacities),
                                ptr(betas), ptr(backgrounds), ptr(masks), ch, int(width), int(height), int(tile_size),
                                ptr(isect_offsets), ptr(flatten_ids), ptr(render_alphas), ptr(last_ids),
                                ptr(v_rc), ptr(v_ra), ptr(v_means2d), ptr(v_conics),
                                ptr(v_colors), ptr(v_opacities), ptr(v_betas), _stream()),
          "ubs_rasterize_bwd")
    return v_means2d, v_conics, v_colors, v_opacities, v_betas


class _RasterizeToPixels(torch.autograd.Function):
    """Mirror of the reference's _RasterizeToPixels (cuda/_wrapper.py:928-1050)."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size,
                isect_offsets, flatten_ids):
        render_colors, render_alphas, last_ids = rasterize_fwd(
            means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size, isect_offsets,
            flatten_ids)
        ctx.save_for_backward(means2d, conics, colors, opacities, betas, backgrounds, masks, isect_offsets,
                              flatten_ids, render_alphas, last_ids)
        ctx.dims = (width, height, tile_size)
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas):
        (means2d, conics, colors, opacities, betas, backgrounds, masks, isect_offsets, flatten_ids, render_alphas,
         last_ids) = ctx.saved_tensors
        width, height, tile_size = ctx.dims
        v_means2d, v_conics, v_colors, v_opacities, v_betas = rasterize_bwd(
            means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size, isect_offsets,
            flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas)
        v_backgrounds = None
        if ctx.needs_input_grad[5]:
            v_backgrounds = (v_render_colors * (1.0 - render_alphas).float()).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_colors, v_opacities, v_betas, v_backgrounds, None, None, None, None, None,
                None)


def rasterize_to_pixels(
    means2d: Tensor,  # [C, N, 2]
    conics: Tensor,  # [C, N, 3]
    colors: Tensor,  # [C, N, channels]
    opacities: Tensor,  # [C, N]
    betas: Tensor,  # [C, N]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C, tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
    backgrounds: Optional[Tensor] = None,  # [C, channels]
    masks: Optional[Tensor] = None,  # [C, tile_height, tile_width]
) -> Tuple[Tensor, Tensor]:
    """Same call as the reference (cuda/_wrapper.py:366-491), including the channel padding rule."""
    C = isect_offsets.size(0)
    device = means2d.device
    N = means2d.size(1)
    assert means2d.shape == (C, N, 2), means2d.shape
    assert conics.shape == (C, N, 3), conics.shape
    assert colors.shape[:2] == (C, N), colors.shape
    assert opacities.shape == (C, N), opacities.shape
    assert betas.shape == (C, N), betas.shape
    if backgrounds is not None:
        assert backgrounds.shape == (C, colors.shape[-1]), backgrounds.shape
        backgrounds = backgrounds.contiguous()
    if masks is not None:
        assert masks.shape == isect_offsets.shape, masks.shape
        masks = masks.contiguous()

    channels = colors.shape[-1]
    if channels > 513 or channels == 0:
        raise ValueError(f"Unsupported number of color channels: {channels}")
    if channels > SUPPORTED_CHANNELS[-1]:
        # channels composite independently: split into launches of <= 16 (the reference instead instantiates its
        # kernel for up to 513 channels, rasterize_to_pixels_fwd.cu:338-357); alphas are identical in every chunk
        w = SUPPORTED_CHANNELS[-1]
        outs, alphas = [], None
        for c0 in range(0, channels, w):
            rc, ra = rasterize_to_pixels(means2d, conics, colors[..., c0:c0 + w], opacities, betas, image_width,
                                         image_height, tile_size, isect_offsets, flatten_ids,
                                         backgrounds[..., c0:c0 + w] if backgrounds is not None else None, masks)
            outs.append(rc)
            alphas = ra if alphas is None else alphas
        return torch.cat(outs, dim=-1), alphas
    padded_channels = 0
    if channels not in SUPPORTED_CHANNELS:
        padded_channels = min(c for c in SUPPORTED_CHANNELS if c >= channels) - channels
        colors = torch.cat([colors, torch.zeros(*colors.shape[:-1], padded_channels, device=device)], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat(
                [backgrounds, torch.zeros(*backgrounds.shape[:-1], padded_channels, device=device)], dim=-1)

    tile_height, tile_width = isect_offsets.shape[1:3]
    assert tile_height * tile_size >= image_height, f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert tile_width * tile_size >= image_width, f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"

    masks_u8 = None if masks is None else masks.to(torch.bool).contiguous().view(torch.uint8)
    render_colors, render_alphas = _RasterizeToPixels.apply(
        _req(means2d, "means2d"), _req(conics, "conics"), _req(colors, "colors"), _req(opacities, "opacities"),
        _req(betas, "betas"), backgrounds, masks_u8, image_width, image_height, tile_size,
        _req(isect_offsets, "isect_offsets", torch.int32), _req(flatten_ids, "flatten_ids", torch.int32))
    if padded_channels > 0:
        render_colors = render_colors[..., :-padded_channels]
    return render_colors, render_alphas


# --------------------------------------------------------------------------------------------------------------
# K1-K4: covariance build and conditioning (companion operators of scene/beta_model.py:12-22)
# --------------------------------------------------------------------------------------------------------------
class _LTriangleToRotmat(torch.autograd.Function):
    """Mirror of cuda/_wrapper.py:614-630."""

    @staticmethod
    def forward(ctx, l_triangle):
        lib = _lib.load()
        N = l_triangle.shape[0]
        rot = torch.empty((N, 3, 3), dtype=torch.float32, device=l_triangle.device)
        check(lib.ubs_l_triangle_to_rotmat_fwd(N, ptr(l_triangle), ptr(rot), _stream()), "ubs_l_triangle_to_rotmat_fwd")
        return rot

    @staticmethod
    def backward(ctx, v_rot):
        lib = _lib.load()
        v_rot = _req(v_rot, "v_rot")
        N = v_rot.shape[0]
        v_lt = torch.empty((N, 3), dtype=torch.float32, device=v_rot.device)
        check(lib.ubs_l_triangle_to_rotmat_bwd(N, ptr(v_rot), ptr(v_lt), _stream()), "ubs_l_triangle_to_rotmat_bwd")
        return v_lt


@_storage_types
def l_triangle_to_rotmat(l_triangle: Tensor) -> Tensor:
    """[N,3] skew parameters -> [N,3,3] first-order rotation I + A (cuda/_wrapper.py:34-36)."""
    assert l_triangle.shape[1] == 3, l_triangle.shape
    return _LTriangleToRotmat.apply(_req(l_triangle, "l_triangle"))


def _check_rest_layout(D: int, rest_i: Tensor, rest_j: Tensor):
    """The kernels hard-wire torch.tril_indices(D, D, -1) order; reject anything else loudly."""
    ti, tj = torch.tril_indices(D, D, offset=-1)
    m = (ti >= 3) | (tj >= 3)
    if rest_i.numel() != int(m.sum()) or not torch.equal(rest_i.cpu().long(), ti[m]) or not torch.equal(
            rest_j.cpu().long(), tj[m]):
        raise NotImplementedError("rest_i/rest_j must be the tril_indices layout built by scene/beta_model.py:69-73")



class _RotScaleLTriangleToCovar(torch.autograd.Function):
    """Mirror of cuda/_wrapper.py:633-684."""

    @staticmethod
    def forward(ctx, rot, scale, l_triangle, spatial_block):
        lib = _lib.load()
        N, D = scale.shape
        d = 3 if (spatial_block or D == 3) else D
        covar = torch.empty((N, d, d), dtype=torch.float32, device=scale.device)
        check(lib.ubs_rot_scale_l_triangle_to_covar_fwd(N, D, 1 if d == 3 else 0, ptr(rot), ptr(scale),
                                                        ptr(l_triangle), ptr(covar), _stream()),
              "ubs_rot_scale_l_triangle_to_covar_fwd")
        ctx.save_for_backward(rot, scale, l_triangle)
        ctx.spatial = d == 3
        return covar

    @staticmethod
    def backward(ctx, v_covar):
        rot, scale, l_triangle = ctx.saved_tensors
        if not any(ctx.needs_input_grad[:3]):
            return None, None, None, None
        lib = _lib.load()
        N, D = scale.shape
        v_rot, v_scale, v_lt = torch.empty_like(rot), torch.empty_like(scale), torch.empty_like(l_triangle)
        v_covar = _req(v_covar, "v_covar")
        check(lib.ubs_rot_scale_l_triangle_to_covar_bwd(N, D, 1 if ctx.spatial else 0, ptr(rot), ptr(scale),
                                                        ptr(l_triangle), ptr(v_covar), ptr(v_rot),
                                                        ptr(v_scale), ptr(v_lt), _stream()),
              "ubs_rot_scale_l_triangle_to_covar_bwd")
        return v_rot, v_scale, v_lt, None


@_storage_types
def rot_scale_l_triangle_to_covar(rot: Tensor, scale: Tensor, l_triangle: Tensor, rest_i: Tensor, rest_j: Tensor,
                                  spatial_block: bool = False) -> Tensor:
    """Sigma = L L^T (cuda/_wrapper.py:39-55).  D in [4, 8]."""
    D = scale.shape[1]
    # validated once per index-tensor OBJECT and content version (needs a host copy); the mark lives on the tensors
    # themselves, so a recycled address can never pass for a validated layout
    mark_i, mark_j = (D, "i", rest_i._version), (D, "j", rest_j._version)
    if getattr(rest_i, "_ubs_layout_ok", None) != mark_i or getattr(rest_j, "_ubs_layout_ok", None) != mark_j:
        _check_rest_layout(D, rest_i, rest_j)
        rest_i._ubs_layout_ok, rest_j._ubs_layout_ok = mark_i, mark_j
    assert rot.shape[1:] == (3, 3) and l_triangle.shape[1] == D * (D - 1) // 2
    return _RotScaleLTriangleToCovar.apply(_req(rot, "rot"), _req(scale, "scale"), _req(l_triangle, "l_triangle"),
                                           bool(spatial_block))


class _CondMeanConvarianceOpacity(torch.autograd.Function):
    """Mirror of cuda/_wrapper.py:573-611 (no gradient for `query`)."""

    @staticmethod
    def forward(ctx, means, covars, opacities, betas, query):
        lib = _lib.load()
        N, D = means.shape
        dev = means.device
        om = torch.empty((N, 3), dtype=torch.float32, device=dev)
        oc = torch.empty((N, 3, 3), dtype=torch.float32, device=dev)
        oo = torch.empty((N, 1), dtype=torch.float32, device=dev)
        check(lib.ubs_cond_mean_covar_opacity_fwd(N, D, ptr(means), ptr(covars), ptr(opacities), ptr(betas),
                                                  ptr(query), ptr(om), ptr(oc), ptr(oo), _stream()),
              "ubs_cond_mean_covar_opacity_fwd")
        ctx.save_for_backward(means, covars, opacities, betas, query)
        return om, oc, oo

    @staticmethod
    def backward(ctx, v_om, v_oc, v_oo):
        means, covars, opacities, betas, query = ctx.saved_tensors
        lib = _lib.load()
        N, D = means.shape
        v_means, v_covars = torch.empty_like(means), torch.empty_like(covars)
        v_opac, v_betas = torch.empty_like(opacities), torch.empty_like(betas)
        v_om, v_oc, v_oo = _req(v_om, "v_means"), _req(v_oc, "v_covars"), _req(v_oo, "v_opacities")  # outlive the launch
        check(lib.ubs_cond_mean_covar_opacity_bwd(N, D, ptr(means), ptr(covars), ptr(opacities), ptr(betas),
                                                  ptr(query), ptr(v_om), ptr(v_oc),
                                                  ptr(v_oo), ptr(v_means), ptr(v_covars),
                                                  ptr(v_opac), ptr(v_betas), _stream()),
              "ubs_cond_mean_covar_opacity_bwd")
        return v_means, v_covars, v_opac, v_betas, None


@_storage_types
def cond_mean_convariance_opacity(means: Tensor, covars: Tensor, opacities: Tensor, betas: Tensor,
                                  query: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Condition the D-dim primitive on `query` (cuda/_wrapper.py:18-31).  means [N,D], covars [N,D,D],
    opacities [N,1], betas/query [N,D-3] -> means [N,3], covars [N,3,3], opacities [N,1]."""
    N, D = means.shape
    assert covars.shape == (N, D, D) and opacities.shape == (N, 1)
    assert betas.shape == (N, D - 3) and query.shape == (N, D - 3)
    return _CondMeanConvarianceOpacity.apply(_req(means, "means"), _req(covars, "covars"),
                                             _req(opacities, "opacities"), _req(betas, "betas"), _req(query, "query"))
