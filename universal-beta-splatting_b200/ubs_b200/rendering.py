"""rasterization(): the drop-in entry point behind which the B200 kernels sit.

Signature, kwargs, return values and `meta` keys follow the reference exactly
(submodules/gsplat/rendering.py:17-232), including its quirks: `l_triagnles` / `scales` are ignored when
`covars` is given (the only way scene/beta_model.py:697-711 calls it), "EDepth" is not alpha-normalised
(rendering.py:219 tests for "ED"), and background is zero for the depth-only modes.

Two routes produce the same results:
  * fused (ubs_b200/dropin.py): taken when `means` / `opacities` / `covars` are the still-deferred outputs of this
    package's own cond_mean_convariance_opacity -- i.e. for the literal statements of BetaModel.render / .view
    (scene/beta_model.py:660-722, 724-831).  One pack pass + the fused kernels, no host synchronisation;
  * operator chain: K5 -> tile lists -> K10 as separate operators, for any other input.
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import dropin
from .ops import fully_fused_projection, isect_tiles, rasterize_to_pixels

RENDER_MODES = ["RGB", "Normal", "Diffuse", "Specular", "Depth", "EDepth", "RGB+D", "RGB+ED"]
_DEPTH_ONLY = ("Depth", "EDepth", "Normal")
_WITH_DEPTH = ("RGB+D", "RGB+ED")


def depth_to_normal(depths: Tensor, camtoworlds: Tensor, Ks: Tensor) -> Tensor:
    """z-depth map [..., H, W, 1] -> world-space normals [..., H, W, 3] by central differences of the
    back-projected points; border pixels are zero (reference: submodules/gsplat/utils.py:40-131)."""
    device = depths.device
    height, width = depths.shape[-3:-1]
    x, y = torch.meshgrid(torch.arange(width, device=device), torch.arange(height, device=device), indexing="xy")
    fx, fy = Ks[..., 0, 0], Ks[..., 1, 1]
    cx, cy = Ks[..., 0, 2], Ks[..., 1, 2]
    dirs = torch.stack([(x - cx[..., None, None] + 0.5) / fx[..., None, None],
                        (y - cy[..., None, None] + 0.5) / fy[..., None, None]], dim=-1)
    dirs = F.pad(dirs, (0, 1), value=1.0)
    dirs = torch.einsum("...ij,...hwj->...hwi", camtoworlds[..., :3, :3], dirs)
    points = camtoworlds[..., :3, -1][..., None, None, :] + depths * dirs
    dx = points[..., 2:, 1:-1, :] - points[..., :-2, 1:-1, :]
    dy = points[..., 1:-1, 2:, :] - points[..., 1:-1, :-2, :]
    normals = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return F.pad(normals, (0, 0, 1, 1, 1, 1), value=0.0)


def _finish(render_colors: Tensor, render_alphas: Tensor, render_mode: str, viewmats: Tensor, Ks: Tensor) -> Tensor:
    """What the reference does to the composited image after K10 (rendering.py:219-230): expected depth for
    "RGB+ED" only (the test there is `in ["ED", "RGB+ED"]`, so "EDepth" stays accumulated depth), normals from the
    depth image for "Normal"."""
    if render_mode in ("ED", "RGB+ED"):
        expected = render_colors[..., -1:] / render_alphas.clamp(min=1e-10)
        render_colors = torch.cat([render_colors[..., :-1], expected], dim=-1)
    if render_mode == "Normal":
        render_colors = (depth_to_normal(render_colors, torch.inverse(viewmats), Ks) + 1) / 2
    return render_colors


def rasterization(
    means: Tensor,  # [N, 3]
    l_triagnles: Tensor,  # ignored when covars is given
    scales: Tensor,  # ignored when covars is given
    opacities: Tensor,  # [N]
    betas: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    assert render_mode in RENDER_MODES, render_mode
    C = viewmats.shape[0]
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    if covars is None:
        # The reference would need [N,4] quaternion-like `l_triagnles` here, a branch UBS never takes.
        raise NotImplementedError("rasterization() requires `covars` (the only form the UBS caller uses)")

    routed = dropin.try_fused_rasterization(means, opacities, betas, colors, viewmats, Ks, width, height, near_plane,
                                            far_plane, radius_clip, eps2d, tile_size, backgrounds, render_mode,
                                            rasterize_mode, covars)
    if routed is not None:
        render_colors, render_alphas, meta = routed
        return _finish(render_colors, render_alphas, render_mode, viewmats, Ks), render_alphas, meta

    # ---- operator chain ------------------------------------------------------------------------------------------
    means, opacities, betas, colors, covars = dropin.materialize((means, opacities, betas, colors, covars))
    if any(isinstance(t, dropin.Deferred) for t in (l_triagnles, scales)):
        l_triagnles, scales = None, None  # unused with covars; never force a computation for them
    dropin._STATS["fallback"] += 1
    N = means.shape[0]
    assert means.shape == (N, 3), means.shape
    assert covars.shape == (N, 3, 3), covars.shape
    assert opacities.shape == (N,), opacities.shape
    assert betas.shape == (N,), betas.shape
    per_camera_colors = colors.dim() == 3
    assert (colors.shape[:2] == (C, N)) if per_camera_colors else (colors.dim() == 2 and colors.shape[0] == N), colors.shape

    upper = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    antialiased = rasterize_mode == "antialiased"
    radii, means2d, depths, conics, compensations = fully_fused_projection(
        means, covars[..., upper[0], upper[1]], None, None, viewmats, Ks, width, height, eps2d=eps2d,
        near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip, sparse_grad=False,
        calc_compensations=antialiased, ortho=False)
    opacities = opacities.repeat(C, 1)  # [C, N]
    betas = betas.repeat(C, 1)
    if compensations is not None:
        opacities = opacities * compensations

    if not per_camera_colors:
        colors = colors.expand(C, -1, -1)
    zero_bg = None if backgrounds is None else torch.zeros(C, 1, device=backgrounds.device)
    if render_mode in _WITH_DEPTH:
        colors = torch.cat((colors, depths[..., None]), dim=-1)
        backgrounds = None if backgrounds is None else torch.cat([backgrounds, zero_bg], dim=-1)
    elif render_mode in _DEPTH_ONLY:
        colors, backgrounds = depths[..., None], zero_bg

    tile_width, tile_height = math.ceil(width / float(tile_size)), math.ceil(height / float(tile_size))
    # tile binning + per-tile depth sort (bit-identical lists, ~2x faster than emit + global radix sort); it needs
    # non-negative depths, which a positive near plane guarantees for every visible primitive
    tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = isect_tiles(
        means2d, radii, depths, tile_size, tile_width, tile_height, n_cameras=C, return_offsets=True,
        method="bin" if near_plane > 0 else "onesweep")
    meta = {"camera_ids": None, "primitive_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
            "conics": conics, "opacities": opacities, "betas": betas, "tile_width": tile_width,
            "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids,
            "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width, "height": height,
            "tile_size": tile_size, "n_cameras": C}

    pieces, render_alphas = [], None
    for c0 in range(0, colors.shape[-1], channel_chunk):
        sl = slice(c0, c0 + channel_chunk)
        rc, ra = rasterize_to_pixels(means2d, conics, colors[..., sl], opacities, betas, width, height, tile_size,
                                     isect_offsets, flatten_ids,
                                     backgrounds=None if backgrounds is None else backgrounds[..., sl])
        pieces.append(rc)
        render_alphas = ra if render_alphas is None else render_alphas
    render_colors = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=-1)
    return _finish(render_colors, render_alphas, render_mode, viewmats, Ks), render_alphas, meta
