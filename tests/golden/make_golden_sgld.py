"""Generates tests/golden/sgld_D{6,7}.npz by executing the REFERENCE's own position-noise statements.

train.py:156-163 is inline code of training(); this script takes exactly those statements (the assignment of
`xyz_covariance`, the two assignments of `noise` and the `beta_model._xyz.add_(noise)` call) out of
/root/reference/train.py with `ast`, unmodified, and executes them against a stub `beta_model` whose
get_xyz_covariance is the reference's own pure-torch K1 / K2 (gsplat/cuda/_torch_impl.py:60-129, spatial_block=True --
the CUDA ops scene/beta_model.py:143-152 calls cannot run here) on the activations of scene/beta_model.py:36-52.
The N(0,1) draw of torch.randn_like is recorded.  The fixture holds tensors only.

    python tests/golden/make_golden_sgld.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/submodules")
from gsplat.cuda import _torch_impl as T  # noqa: E402

SRC = "/root/reference/train.py"


def reference_statements():
    text = open(SRC).read()
    lines = text.splitlines()
    picked = []
    for node in ast.walk(ast.parse(text)):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in ("xyz_covariance", "noise"):
            picked.append(node)
        if isinstance(node, ast.Expr) and "beta_model._xyz.add_(noise)" in ast.unparse(node):
            picked.append(node)
    picked.sort(key=lambda n: n.lineno)
    assert len(picked) == 4, [ast.unparse(n) for n in picked]
    import textwrap
    return "\n".join(textwrap.dedent("\n".join(lines[n.lineno - 1:n.end_lineno])) for n in picked)


class StubModel:
    """What the statements touch of BetaModel: _xyz, get_opacity, get_xyz_covariance."""

    def __init__(self, D, N, seed):
        g = torch.Generator().manual_seed(seed)
        self.D = D
        self._xyz = torch.randn(N, 3, generator=g)
        self._opacity = torch.randn(N, 1, generator=g) * 2.0 - 3.0  # mostly transparent: (1 - o)^100 not all zero
        self._scale = torch.randn(N, D, generator=g) * 0.5 - 1.0
        self._l_triangle = torch.randn(N, D * (D - 1) // 2, generator=g) * 0.3
        ti, tj = torch.tril_indices(D, D, offset=-1)
        m = (ti >= 3) | (tj >= 3)
        self.rest_i, self.rest_j = ti[m].to(torch.int32), tj[m].to(torch.int32)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_xyz_covariance(self):
        rot = T._l_triangle_to_rotmat(self._l_triangle[:, :3])
        return T._rot_scale_l_triangle_to_covar(rot, torch.nn.functional.softplus(self._scale), self._l_triangle,
                                                self.rest_i, self.rest_j, spatial_block=True)


if __name__ == "__main__":
    code = reference_statements()
    print(code)
    for D in (6, 7):
        bm = StubModel(D, 300, 31 + D)
        out = {"xyz": bm._xyz.numpy().copy(), "opacity": bm._opacity.numpy().copy(), "scale": bm._scale.numpy().copy(),
               "l_triangle": bm._l_triangle.numpy().copy()}
        drawn = []
        real_randn_like = torch.randn_like

        def recording_randn_like(t, *a, **k):
            r = real_randn_like(t, *a, **k)
            drawn.append(r.clone())
            return r

        torch_proxy = types.SimpleNamespace(**{k: getattr(torch, k) for k in ("pow", "bmm")}, randn_like=recording_randn_like)
        args = types.SimpleNamespace(noise_lr=1.0)  # arguments/__init__.py:99
        xyz_lr = 1.6e-4 * 0.73
        torch.manual_seed(500 + D)
        with torch.no_grad():
            exec(code, {"torch": torch_proxy, "beta_model": bm, "args": args, "xyz_lr": xyz_lr})
        out.update(noise=drawn[0].numpy(), noise_lr=np.float64(args.noise_lr), xyz_lr=np.float64(xyz_lr),
                   xyz_out=bm._xyz.numpy().copy())
        np.savez_compressed(os.path.join(HERE, f"sgld_D{D}.npz"), **out)
        print(D, "max displacement", float(np.abs(out["xyz_out"] - out["xyz"]).max()))
