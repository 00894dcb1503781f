"""Host-side mirror of the renderer-facing part of BetaModel (scene/beta_model.py) over the packed records.

  * PackedBetaModel.render(camera, render_mode, mask)   <->  BetaModel.render      (scene/beta_model.py:660-722)
  * PackedBetaModel.view(c2w, K, W, H, ...)              <->  BetaModel.view        (scene/beta_model.py:724-831)
  * quantile_mask                                        <->  the viewer's beta-quantile primitive filter (:729-755)

Every render mode goes through the fused fast path (fused.FusedRasterizer; the primitive mask is applied inside the
projection kernel, no gather of the parameters): "RGB" / the viewer's "Alpha" composite the RGB of the 48-byte splat
rows, "RGB+D" / "RGB+ED" its RGB + depth, "Depth" / "EDepth" / "Normal" its depth alone
(submodules/gsplat/rendering.py:131-142,219-230).  `chain=True` forces the reference-shaped operator chain
(ops.* + rendering.rasterization) instead -- what the parity tests compare the fused modes with.  The colour mapping
of depth images (utils/general_utils.py:196-220, matplotlib's "turbo") is presentation and stays with the viewer.
"""
from typing import Dict, Optional, Sequence

import torch
from torch import Tensor

from . import fused, ops
from .rendering import _finish, rasterization

_CHANNELS = {"RGB": 3, "Alpha": 3, "RGB+D": 4, "RGB+ED": 4, "Depth": 1, "EDepth": 1, "Normal": 1}


def quantile_mask(beta_raw: Tensor, b_xyz=(0, 100), b_view=(0, 100), b_time: Optional[Sequence[float]] = (0, 100)):
    """scene/beta_model.py:729-755: keep primitives whose raw spatial / mean view / time beta lies between the
    given percentiles.  beta_raw: [N, D-2] (column 0 spatial, 1:4 view, 4 time)."""
    x = beta_raw[:, 0]
    v = beta_raw[:, 1:4].mean(dim=-1)
    mask = ((x >= x.quantile(b_xyz[0] / 100)) & (x <= x.quantile(b_xyz[1] / 100))
            & (v >= v.quantile(b_view[0] / 100)) & (v <= v.quantile(b_view[1] / 100)))
    if b_time is not None:
        t = beta_raw[:, 4]
        mask = mask & (t >= t.quantile(b_time[0] / 100)) & (t <= t.quantile(b_time[1] / 100))
    return mask


class PackedBetaModel:
    """A trained / training model as ONE packed [N, stride] record buffer plus the renderer state around it."""

    def __init__(self, records: Tensor, D: int, background: Optional[Tensor] = None):
        assert records.is_cuda and records.shape[1] == fused.record_stride(D)
        self.records, self.D = records, D
        self.background = torch.zeros(3, device=records.device) if background is None else background
        self._rz: Dict[tuple, fused.FusedRasterizer] = {}

    @property
    def N(self) -> int:
        return self.records.shape[0]

    # ---- the seven tensors and their activations (scene/beta_model.py:36-52,103-125) ---------------------------
    def tensors(self):
        return fused.unpack_records(self.D, self.records)

    def _rasterizer(self, W, H, near, far, clip) -> fused.FusedRasterizer:
        key = (self.N, W, H, near, far, clip)
        rz = self._rz.get(key)
        if rz is None:
            if len(self._rz) >= 4:  # a viewer changes its window size: keep the cache bounded
                self._rz.pop(next(iter(self._rz)))
            rz = fused.FusedRasterizer(self.D, self.N, W, H, n_cams=1, device=self.records.device, near_plane=near,
                                       far_plane=far, radius_clip=clip)
            self._rz[key] = rz
        return rz

    def _conditioned(self, cam_pos: Tensor, timestamp: float):
        """K1 -> K2 -> K3 with the stand-alone operators, as get_cond_mean_convariance_opacity does (:154-159)."""
        xyz, mean, rgb, opacity, beta, scale, l_tri = [t.contiguous() for t in self.tensors()]
        D = self.D
        s = torch.nn.functional.softplus(scale)
        o = torch.sigmoid(opacity)
        b = 4.0 * torch.exp(beta)
        ti, tj = torch.tril_indices(D, D, offset=-1)
        m = (ti >= 3) | (tj >= 3)
        rest_i, rest_j = ti[m].to(torch.int32).to(xyz.device), tj[m].to(torch.int32).to(xyz.device)
        rot = ops.l_triangle_to_rotmat(l_tri[:, :3].contiguous())
        covar = ops.rot_scale_l_triangle_to_covar(rot, s, l_tri, rest_i, rest_j)
        view_dir = xyz - cam_pos.unsqueeze(0)
        view_dir = view_dir / view_dir.norm(dim=-1, keepdim=True)
        if D == 6:
            q = view_dir
        elif D == 7:
            q = torch.cat([view_dir, torch.full((xyz.shape[0], 1), float(timestamp), device=xyz.device)], dim=-1)
        else:
            raise NotImplementedError("Only implemented for 6D or 7D query")
        means, covs, opac = ops.cond_mean_convariance_opacity(torch.cat([xyz, mean], dim=-1), covar, o,
                                                              b[:, 1:].contiguous(), q.contiguous())
        return means, covs, opac.squeeze(-1), b[:, 0].contiguous(), rgb

    @torch.no_grad()
    def _render(self, viewmat, K, cam_pos, timestamp, W, H, render_mode, mask, near, far, clip, chain=False):
        dev = self.records.device
        bg = self.background.reshape(1, 3).to(dev)
        if not chain:
            ch = _CHANNELS[render_mode]
            rz = self._rasterizer(W, H, near, far, clip)
            ts = torch.tensor([float(timestamp)], device=dev) if self.D == 7 else None
            # depth channel / depth-only images composite over a zero background (rendering.py:131-142)
            bg_ch = bg if ch == 3 else (torch.cat([bg, bg.new_zeros(1, 1)], dim=-1) if ch == 4 else bg.new_zeros(1, 1))
            vm, Kc = viewmat[None].contiguous(), K[None].contiguous()
            rc, ra = rz.forward(self.records, vm, Kc, cam_pos[None].contiguous(), ts, bg_ch.contiguous(),
                                prim_mask=mask, channels=ch)
            rc = _finish(rc, ra, "RGB" if render_mode == "Alpha" else render_mode, vm, Kc)
            meta = {"means2d": rz.means2d, "radii": rz.radii}
            return rc, ra, meta
        means, covs, opac, beta0, rgb = self._conditioned(cam_pos, timestamp)
        if mask is None:
            mask = torch.ones(self.N, dtype=torch.bool, device=dev)
        return rasterization(means[mask], None, None, opac[mask], beta0[mask], rgb[mask], viewmat[None], K[None], W, H,
                             near_plane=near, far_plane=far, radius_clip=clip, backgrounds=bg,
                             render_mode="RGB" if render_mode == "Alpha" else render_mode, covars=covs[mask])

    def render(self, camera, render_mode: str = "RGB", mask: Optional[Tensor] = None, chain: bool = False):
        """camera: anything with viewmat [4,4] (world->camera, row-major), K [3,3], cam_pos [3], width, height,
        timestamp (synth.Camera).  Returns the reference's dict (scene/beta_model.py:716-722); with a mask, the
        fused path keeps the per-primitive outputs at full length N (masked-out primitives have radius 0), the
        operator chain returns them for the kept primitives only, like the reference."""
        rc, ra, meta = self._render(camera.viewmat, camera.K, camera.cam_pos, camera.timestamp, camera.width,
                                    camera.height, render_mode, mask, 0.01, 1e10, 0.0, chain)
        return {"render": rc.permute(0, 3, 1, 2).contiguous()[0], "alpha": ra, "viewspace_points": meta["means2d"],
                "visibility_filter": meta["radii"] > 0, "radii": meta["radii"], "is_used": meta["radii"] > 0}

    @torch.no_grad()
    def view(self, c2w: Tensor, K: Tensor, W: int, H: int, render_mode: str = "RGB", b_xyz=(0, 100), b_view=(0, 100),
             b_time=(0, 100), timestamp: float = 0.0, near_plane: float = 0.01, far_plane: float = 1e10,
             radius_clip: float = 0.0, backgrounds=(0, 0, 0), chain: bool = False):
        """The viewer callback (scene/beta_model.py:724-831) without the GUI state object: returns
        (image [H,W,ch] on the device, rendered primitive count).  Depth modes return the raw 1-channel image."""
        dev = self.records.device
        beta_raw = self.tensors()[4]
        full = tuple(b_xyz) == (0, 100) and tuple(b_view) == (0, 100) and (self.D != 7 or tuple(b_time) == (0, 100))
        mask = None if full else quantile_mask(beta_raw, b_xyz, b_view, b_time if self.D == 7 else None)
        self.background = torch.tensor(backgrounds, device=dev, dtype=torch.float32) / 255.0
        viewmat = torch.linalg.inv(c2w)
        rc, ra, meta = self._render(viewmat, K, c2w[:3, 3].contiguous(), timestamp, W, H,
                                    render_mode, mask, near_plane, far_plane, radius_clip, chain)
        if render_mode == "Alpha":
            rc = ra
        return rc[0], int((meta["radii"] > 0).sum().item())
