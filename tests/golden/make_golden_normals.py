"""Generates tests/golden/depth_to_normal.npz from the reference's own `depth_to_normal`
(submodules/gsplat/utils.py:40-131), imported in the build container (CPU).  Run from the repo root:

    python tests/golden/make_golden_normals.py
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/submodules/gsplat/utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "depth_to_normal.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_gsplat_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(20251003)
    C, H, W = 2, 37, 53
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    depths = (3.0 + torch.sin(4 * xx) * 0.5 + yy + 0.05 * torch.randn(C, H, W, generator=g))[..., None].float()
    c2w = torch.eye(4).repeat(C, 1, 1)
    a = torch.linalg.qr(torch.randn(C, 3, 3, generator=g))[0]
    c2w[:, :3, :3] = a
    c2w[:, :3, 3] = torch.randn(C, 3, generator=g)
    Ks = torch.tensor([[60.0, 0, W / 2], [0, 55.0, H / 2], [0, 0, 1]]).repeat(C, 1, 1)
    normals = mod.depth_to_normal(depths, c2w, Ks)
    np.savez_compressed(OUT, depths=depths.numpy(), c2w=c2w.numpy(), Ks=Ks.numpy(), normals=normals.numpy())
    print("wrote", OUT, normals.shape)


if __name__ == "__main__":
    main()
