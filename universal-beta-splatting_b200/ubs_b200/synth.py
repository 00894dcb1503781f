"""Deterministic synthetic scenes and cameras for the BASELINE.json configs (generators per SURVEY.md 8(d)).

Everything is generated with a seeded torch.Generator on the CPU and moved to the requested device, so the GPU
tests, the CPU oracle and the benchmarks all see identical inputs.  Parameter tensors use the reference's raw
(pre-activation) conventions (scene/beta_model.py:36-52,161-228).
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch

BASE_SEED = 20251003


@dataclass
class Scene:
    """The 7 raw parameter tensors of a BetaModel (scene/beta_model.py:57-63)."""
    D: int
    xyz: torch.Tensor  # [N,3]
    mean: torch.Tensor  # [N,D-3]
    rgb: torch.Tensor  # [N,3]
    opacity: torch.Tensor  # [N,1] logit
    beta: torch.Tensor  # [N,D-2] raw (activation 4*exp)
    scale: torch.Tensor  # [N,D] raw (activation softplus)
    l_triangle: torch.Tensor  # [N,D(D-1)/2]

    @property
    def N(self):
        return self.xyz.shape[0]

    def to(self, device):
        return Scene(self.D, *[t.to(device) for t in (self.xyz, self.mean, self.rgb, self.opacity, self.beta,
                                                       self.scale, self.l_triangle)])

    def tensors(self):
        return [self.xyz, self.mean, self.rgb, self.opacity, self.beta, self.scale, self.l_triangle]


@dataclass
class Camera:
    viewmat: torch.Tensor  # [4,4] world->camera (row-major)
    K: torch.Tensor  # [3,3]
    cam_pos: torch.Tensor  # [3]
    width: int
    height: int
    timestamp: float = 0.0


def inverse_softplus(y):
    return y + torch.log(-torch.expm1(-y))


def make_scene(N: int, D: int = 6, seed: int = BASE_SEED, unbounded: bool = False, extent: float = 4.0,
               device="cpu") -> Scene:
    g = torch.Generator().manual_seed(seed)
    U = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    Nn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    xyz = (U(N, 3) * 2 - 1) * extent
    if unbounded:
        n_far = N // 5
        d = Nn(n_far, 3)
        d = d / d.norm(dim=-1, keepdim=True)
        xyz[:n_far] = d * (8 + 32 * U(n_far, 1))
    mean = U(N, D - 3) * 2 - 1
    if D == 7:
        mean[:, 3] = U(N)
    s_spatial = torch.exp(math.log(0.005) + U(N, 3) * (math.log(0.05) - math.log(0.005))) * extent / (N / 1e5) ** (1 / 3)
    s_cond = 0.5 + 1.5 * U(N, D - 3)
    scale = inverse_softplus(torch.cat([s_spatial, s_cond], dim=-1))
    l_triangle = Nn(N, D * (D - 1) // 2) * 0.05
    opacity = Nn(N, 1) * 1.5
    beta = Nn(N, D - 2) * 0.5
    if D == 7:
        beta[:, 1:4] -= 3.0
    rgb = U(N, 3)
    return Scene(D, xyz, mean, rgb, opacity, beta, scale, l_triangle).to(device)


def look_at(eye: torch.Tensor, target: torch.Tensor, up=(0.0, 0.0, 1.0)) -> torch.Tensor:
    """World->camera matrix with +z forward, +x right, +y down (the gsplat/COLMAP convention)."""
    up = torch.tensor(up, dtype=torch.float32)
    f = target - eye
    f = f / f.norm()
    r = torch.linalg.cross(f, up)
    r = r / r.norm()
    d = torch.linalg.cross(f, r)
    R = torch.stack([r, d, f], dim=0)
    V = torch.eye(4)
    V[:3, :3] = R
    V[:3, 3] = -R @ eye
    return V


def make_cameras(n: int, width: int, height: int, radius: float = 8.0, fov_y_deg: float = 50.0,
                 seed: int = BASE_SEED, timestamps: Optional[list] = None, device="cpu"):
    g = torch.Generator().manual_seed(seed + 7919)
    cams = []
    fy = 0.5 * height / math.tan(math.radians(fov_y_deg) / 2)
    fx = fy
    K = torch.tensor([[fx, 0.0, width / 2], [0.0, fy, height / 2], [0.0, 0.0, 1.0]])
    for k in range(n):
        ang = 2 * math.pi * k / max(n, 1) + 0.1
        h = -1.0 + 3.0 * torch.rand(1, generator=g).item()
        eye = torch.tensor([radius * math.cos(ang), radius * math.sin(ang), h])
        V = look_at(eye, torch.zeros(3))
        ts = 0.0 if timestamps is None else float(timestamps[k % len(timestamps)])
        cams.append(Camera(V.to(device), K.clone().to(device), eye.to(device), width, height, ts))
    return cams


def random_spd_covars(N: int, seed: int, scale_lo=0.005, scale_hi=0.2, device="cpu"):
    """Random symmetric positive-definite 3x3 covariances [N,3,3] (for stage-isolated projection tests)."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(N, 3, 3, generator=g)
    Q, _ = torch.linalg.qr(A)
    s = torch.exp(math.log(scale_lo) + torch.rand(N, 3, generator=g) * (math.log(scale_hi) - math.log(scale_lo)))
    cov = Q @ torch.diag_embed(s * s) @ Q.transpose(-1, -2)
    cov = 0.5 * (cov + cov.transpose(-1, -2))
    return cov.to(device)


# BASELINE.json configs (index = cfg number - 1); "cams" is the number of cameras in one job.
CONFIGS = {
    "cfg1": dict(N=100_000, D=6, width=800, height=800, cams=1, unbounded=False, radius=8.0, bg=(0.0, 0.0, 0.0)),
    "cfg2": dict(N=300_000, D=6, width=800, height=800, cams=1, unbounded=False, radius=8.0, bg=(1.0, 1.0, 1.0)),
    "cfg3_r4": dict(N=3_000_000, D=6, width=1245, height=825, cams=1, unbounded=True, radius=6.0, bg=(0.0, 0.0, 0.0)),
    "cfg3": dict(N=3_000_000, D=6, width=1920, height=1080, cams=1, unbounded=True, radius=6.0, bg=(0.0, 0.0, 0.0)),
    "cfg4": dict(N=1_000_000, D=7, width=1352, height=1014, cams=300, unbounded=False, radius=8.0, bg=(0.0, 0.0, 0.0)),
    "cfg5": dict(N=3_000_000, D=6, width=1920, height=1080, cams=64, unbounded=True, radius=6.0, bg=(0.0, 0.0, 0.0)),
}


def make_config(name: str, device="cpu", n_override: Optional[int] = None, cams_override: Optional[int] = None):
    cfg = dict(CONFIGS[name])
    idx = list(CONFIGS).index(name)
    N = n_override or cfg["N"]
    n_cams = cams_override or cfg["cams"]
    scene = make_scene(N, cfg["D"], seed=BASE_SEED + idx, unbounded=cfg["unbounded"], device=device)
    ts = [k / 299.0 for k in range(300)] if cfg["D"] == 7 else None
    cams = make_cameras(n_cams, cfg["width"], cfg["height"], radius=cfg["radius"], seed=BASE_SEED + idx,
                        timestamps=ts, device=device)
    bg = torch.tensor(cfg["bg"], dtype=torch.float32, device=device)
    return scene, cams, bg, cfg
