"""Generates tests/golden/torch_impl_D{6,7}.npz by IMPORTING the reference's own pure-PyTorch implementation
(/root/reference/submodules/gsplat/cuda/_torch_impl.py) in the build container.  The reference tree does not exist
on the GPU box, so the outputs are committed as fixtures together with this script.

    python tests/golden/make_golden_torch_impl.py

Covered reference functions: _l_triangle_to_rotmat (:60-88), _rot_scale_l_triangle_to_covar (:94-129),
_cond_mean_convariance_opacity (:9-57), _fully_fused_projection (:307-380, 3-sigma radius),
_isect_tiles (:383-452), _isect_offset_encode (:455-482).
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "universal-beta-splatting_b200"))
sys.path.insert(0, "/root/reference/submodules")

from gsplat.cuda import _torch_impl as T  # noqa: E402
from ubs_b200 import synth  # noqa: E402


def main():
    W, H, TS = 160, 112, 16
    for D in (6, 7):
        N = 400
        sc = synth.make_scene(N, D, seed=4242 + D, extent=2.0)
        # larger footprints than the benchmark scenes so that 400 primitives give multi-tile lists
        cams = synth.make_cameras(2, W, H, radius=6.0, seed=11, timestamps=[0.3, 0.7])
        scale = torch.nn.functional.softplus(sc.scale) * 0.6
        opacity = torch.sigmoid(sc.opacity)
        beta = 4.0 * torch.exp(sc.beta)
        mean = torch.cat([sc.xyz, sc.mean], dim=-1)
        ti, tj = torch.tril_indices(D, D, offset=-1)
        m = (ti >= 3) | (tj >= 3)
        rot = T._l_triangle_to_rotmat(sc.l_triangle[:, :3])
        covar = T._rot_scale_l_triangle_to_covar(rot, scale, sc.l_triangle, ti[m], tj[m], False)
        covar_sp = T._rot_scale_l_triangle_to_covar(rot, scale, sc.l_triangle, ti[m], tj[m], True)
        out = dict(D=D, N=N, W=W, H=H, tile_size=TS, xyz=sc.xyz, mean=mean, scale=scale, opacity=opacity, beta=beta,
                   l_triangle=sc.l_triangle, rgb=sc.rgb, rot=rot, covar=covar, covar_spatial=covar_sp)
        V = torch.stack([c.viewmat for c in cams])
        K = torch.stack([c.K for c in cams])
        out.update(viewmats=V, Ks=K, cam_pos=torch.stack([c.cam_pos for c in cams]),
                   timestamps=torch.tensor([c.timestamp for c in cams]))
        for ci, cam in enumerate(cams):
            vd = sc.xyz - cam.cam_pos[None]
            vd = vd / vd.norm(dim=-1, keepdim=True)
            q = vd if D == 6 else torch.cat([vd, torch.full((N, 1), cam.timestamp)], dim=-1)
            m3, v3, oc = T._cond_mean_convariance_opacity(mean, covar, opacity, beta[:, 1:], q)
            out.update({"query%d" % ci: q, "cond_means%d" % ci: m3, "cond_covars%d" % ci: v3, "cond_opac%d" % ci: oc})
            radii, means2d, depths, conics, comps = T._fully_fused_projection(
                m3, v3, cam.viewmat[None], cam.K[None], W, H, eps2d=0.3, near_plane=0.01, far_plane=1e10,
                calc_compensations=True)
            tw, th = math.ceil(W / TS), math.ceil(H / TS)
            tpg, ids, flat = T._isect_tiles(means2d, radii, depths, TS, tw, th, sort=True)
            off = T._isect_offset_encode(ids, 1, tw, th)
            out.update({"radii%d" % ci: radii, "means2d%d" % ci: means2d, "depths%d" % ci: depths,
                        "conics%d" % ci: conics, "comps%d" % ci: comps, "tiles_per_gauss%d" % ci: tpg,
                        "isect_ids%d" % ci: ids, "flatten_ids%d" % ci: flat, "offsets%d" % ci: off})
            print("D=%d cam%d visible %d pairs %d" % (D, ci, int((radii > 0).sum()), ids.numel()))
        np.savez_compressed(os.path.join(HERE, "torch_impl_D%d.npz" % D),
                            **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
