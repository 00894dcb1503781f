"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs may import this module;
the product (universal-beta-splatting_b200/) never does.

What is restated, and from where (paths relative to /root/reference):
  * activations + query glue ......... scene/beta_model.py:36-52,103-121,660-711
  * K1 l_triangle_to_rotmat .......... submodules/gsplat/cuda/csrc/l_triagnle_to_rotmat_fwd.cu:8-36
  * K2 rot_scale_l_triangle_to_covar . .../rot_scale_l_triangle_to_covar_fwd.cu:8-188
  * K3 cond_mean_convariance_opacity . .../cond_mean_convariance_opacity_fwd.cu:106-290 (CUDA guards, not torch's)
  * K5 fully_fused_projection ........ .../fully_fused_projection_fwd.cu:43-177 + utils.cuh:252-292,374-413,437-466
  * K7-K9, K10, K11 .................. oracle/raster_oracle.c (plain C + OpenMP, via ctypes)
The per-primitive stages are vectorised torch on the CPU, written so that torch.autograd of the restated forward is
the backward oracle for the reference's hand-written VJPs (K4, K6, K2/K1 bwd); `query` is detached because the
reference returns no gradient for it (submodules/gsplat/cuda/_wrapper.py:611).

Parity pinning: the reference has no tests or golden vectors.  This oracle is pinned by
tests/test_oracle_golden.py against (1) tests/golden/torch_impl_D{6,7}.npz -- outputs of the reference's own
`_torch_impl.py` imported in the build container (generator: tests/golden/make_golden_torch_impl.py), with the
documented CUDA-vs-torch divergences (radius extent 1 sigma vs 3 sigma) exposed as the `extent` argument; and
(2) tests/golden/ref_cuda_D{6,7}.npz -- outputs of the reference's compiled CUDA kernels on a B200
(generator: tests/golden/make_golden_ref_cuda.py).
"""
import ctypes
import math
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p, POINTER

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "raster_oracle.c")
_LIB_PATH = os.path.join(_HERE, "_build", "libubs_oracle.so")
_lib = None


def build(force=False):
    """gcc build of oracle/raster_oracle.c -> oracle/_build/libubs_oracle.so (git-ignored)."""
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", _SRC, "-o", _LIB_PATH, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stderr)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_isect_count.restype = c_int64
        _lib.oracle_id_bits.restype = c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _np(t, dtype):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


# ----------------------------------------------------------------------------------------------------------------
# per-primitive stages (torch, CPU, differentiable)
# ----------------------------------------------------------------------------------------------------------------
def activations(xyz, mean, opacity, beta, scale):
    """scene/beta_model.py:36-52,103-121: softplus scale, sigmoid opacity, 4*exp beta, mean = cat(xyz, mean)."""
    return (torch.nn.functional.softplus(scale), torch.sigmoid(opacity), 4.0 * torch.exp(beta),
            torch.cat([xyz, mean], dim=-1))


def l_triangle_to_rotmat(lt3):
    """R = I + skew(a01, a02, a12) (l_triagnle_to_rotmat_fwd.cu:18-35)."""
    a01, a02, a12 = lt3[:, 0], lt3[:, 1], lt3[:, 2]
    one, = (torch.ones_like(a01),)
    return torch.stack([one, a01, a02, -a01, one, a12, -a02, -a12, one], dim=-1).reshape(-1, 3, 3)


def rot_scale_l_triangle_to_covar(rot, scale, l_triangle, spatial_block=False):
    """Sigma = L L^T with L = [[R diag(s0..2), 0], [l_rest, diag(s3..)]] (rot_scale_l_triangle_to_covar_fwd.cu:28-187);
    strictly-lower entries in torch.tril_indices(D, D, -1) order, the layout scene/beta_model.py:69-73 builds."""
    N, D = scale.shape
    L_xyz = rot * scale[:, None, :3]
    if spatial_block or D == 3:
        return L_xyz @ L_xyz.transpose(-1, -2)
    L = torch.zeros((N, D, D), dtype=scale.dtype)
    L[:, :3, :3] = L_xyz
    idx = torch.arange(3, D)
    L[:, idx, idx] = scale[:, 3:]
    ti, tj = torch.tril_indices(D, D, offset=-1)
    m = (ti >= 3) | (tj >= 3)
    L[:, ti[m], tj[m]] = l_triangle[:, 3:]
    return L @ L.transpose(-1, -2)


def _inv3x3_adjugate(a):
    """cond_mean_convariance_opacity_fwd.cu:77-104 (det == 0 -> 1e-20)."""
    a00, a01, a02, a10, a11, a12, a20, a21, a22 = [a[:, i, j] for i in range(3) for j in range(3)]
    c00 = a11 * a22 - a12 * a21
    c01 = -(a10 * a22 - a12 * a20)
    c02 = a10 * a21 - a11 * a20
    c10 = -(a01 * a22 - a02 * a21)
    c11 = a00 * a22 - a02 * a20
    c12 = -(a00 * a21 - a01 * a20)
    c20 = a01 * a12 - a02 * a11
    c21 = -(a00 * a12 - a02 * a10)
    c22 = a00 * a11 - a01 * a10
    det = a00 * c00 + a01 * c01 + a02 * c02
    det = torch.where(det == 0, torch.full_like(det, 1e-20), det)
    inv = torch.stack([c00, c10, c20, c01, c11, c21, c02, c12, c22], dim=-1) / det[:, None]
    return inv.reshape(-1, 3, 3)


def _guard(den):
    return torch.where(den == 0, torch.full_like(den, 1e-20), den)


def _cond_parts(means, covars, opacities, betas, query):
    """Forward of K3 with every intermediate the hand-written backward needs
    (cond_mean_convariance_opacity_fwd.cu:135-289)."""
    N, D = means.shape
    Cd = D - 3
    v11, v12 = covars[:, :3, :3], covars[:, :3, 3:]
    v21, v22 = covars[:, 3:, :3], covars[:, 3:, 3:]
    x = query - means[:, 3:]
    beta_adj = torch.clamp_max(betas * 0.25, 1.0)
    # Gauss-Jordan with partial pivoting for Cd != 3 (fwd.cu:26-75) == LU-based inverse up to rounding
    vinv = _inv3x3_adjugate(v22) if Cd == 3 else torch.linalg.inv(v22)
    r_ = v12 @ vinv
    rb = r_ * beta_adj[:, None, :]
    m_cond = means[:, :3] + (rb @ x[:, :, None])[:, :, 0]
    v_cond = v11 - rb @ v21
    # Cholesky of v22 reading its lower triangle, with the kernel's guards (fwd.cu:250-268)
    Lc = [[None] * Cd for _ in range(Cd)]
    for i in range(Cd):
        for j in range(i + 1):
            s = v22[:, i, j]
            for k in range(j):
                s = s - Lc[i][k] * Lc[j][k]
            if i == j:
                Lc[i][j] = torch.sqrt(torch.where(s <= 0, torch.full_like(s, 1e-20), s))
            else:
                Lc[i][j] = s / _guard(Lc[j][j])
    y = []
    for i in range(Cd):  # forward solve Lc y = x (fwd.cu:270-279)
        s = x[:, i]
        for k in range(i):
            s = s - Lc[i][k] * y[k]
        y.append(s / _guard(Lc[i][i]))
    upper = 1.0 - torch.finfo(torch.float32).eps
    d = [torch.clamp(torch.tanh(y[i] * y[i]), 0.0, upper) for i in range(Cd)]
    o_change = torch.ones_like(y[0])
    for i in range(Cd):
        o_change = o_change * torch.pow(1.0 - d[i], betas[:, i])
    return dict(x=x, beta_adj=beta_adj, vinv=vinv, r_=r_, rb=rb, v12=v12, v21=v21, Lc=Lc, y=y, d=d,
                o_change=o_change, m_cond=m_cond, v_cond=v_cond, upper=upper)


class _Conditioning(torch.autograd.Function):
    """K3 forward + the reference's HAND-WRITTEN backward K4 (cond_mean_convariance_opacity_bwd.cu:101-563).

    The backward is restated line by line rather than left to autograd because the reference is not the exact
    derivative of its forward: in the Cholesky backward it forms G = U L^-1 by solving with L where L^T is needed
    (bwd.cu:495-510), so the opacity-path contribution to gV22 differs from autograd by several percent whenever L
    is not diagonal.  Parity is defined against the reference, so the oracle (and the CUDA library) reproduce it."""

    @staticmethod
    def forward(ctx, means, covars, opacities, betas, query):
        p = _cond_parts(means, covars, opacities, betas, query)
        ctx.save_for_backward(means, covars, opacities, betas, query)
        return p["m_cond"], p["v_cond"], opacities * p["o_change"][:, None]

    @staticmethod
    def backward(ctx, gM, gV, gO):
        means, covars, opacities, betas, query = ctx.saved_tensors
        p = _cond_parts(means, covars, opacities, betas, query)
        N, D = means.shape
        Cd = D - 3
        x, beta_adj, i22, r_, rb, v12, v21 = p["x"], p["beta_adj"], p["vinv"], p["r_"], p["rb"], p["v12"], p["v21"]
        Lc, y, d, o_change, upper = p["Lc"], p["y"], p["d"], p["o_change"], p["upper"]
        gm = torch.zeros_like(means)
        gVfull = torch.zeros_like(covars)
        gm[:, :3] = gM  # bwd.cu:235
        Gr = gM[:, :, None] * x[:, None, :] - gV @ v21.transpose(1, 2)  # bwd.cu:238-266
        gx = torch.einsum("nrc,nr->nc", rb, gM)  # bwd.cu:247-253
        gVfull[:, :3, :3] = gV  # bwd.cu:257-259
        gVfull[:, 3:, :3] = -(rb.transpose(1, 2) @ gV)  # bwd.cu:269-275
        dL_dba = (Gr * r_).sum(dim=1)  # bwd.cu:279-285
        G_r = Gr * beta_adj[:, None, :]
        gVfull[:, :3, 3:] = G_r @ i22.transpose(1, 2)  # bwd.cu:296-304
        Gi22 = v12.transpose(1, 2) @ G_r  # bwd.cu:307-316
        # ---- opacity path (bwd.cu:318-527)
        opa, gOs = opacities[:, 0], gO[:, 0]
        go = (gOs * o_change)[:, None]
        g_o_change = gOs * opa
        gb = torch.zeros_like(betas)
        g_y = []
        for i in range(Cd):
            base = 1.0 - d[i]
            pos = base > 0
            safe = torch.where(pos, base, torch.ones_like(base))
            gb[:, i] = torch.where(pos, g_o_change * o_change * torch.log(safe), torch.zeros_like(base))
            active = pos & (d[i] > 0) & (d[i] < upper)
            g_d = torch.where(active, g_o_change * (-o_change * betas[:, i] / safe), torch.zeros_like(base))
            g_y.append(g_d * (2.0 * y[i] * (1.0 - d[i] * d[i])))
        a = [None] * Cd
        for i in range(Cd - 1, -1, -1):  # (L^T) a = g_y, bwd.cu:414-421
            s = g_y[i]
            for k in range(i + 1, Cd):
                s = s - Lc[k][i] * a[k]
            a[i] = s / _guard(Lc[i][i])
        gx = gx + torch.stack(a, dim=1)
        zero = torch.zeros_like(y[0])
        L = [[Lc[r][c] if c <= r else zero for c in range(Cd)] for r in range(Cd)]
        gL = [[-(a[r] * y[c]) if c <= r else zero for c in range(Cd)] for r in range(Cd)]
        S = [[zero for _ in range(Cd)] for _ in range(Cd)]
        for r in range(Cd):  # S = tril(L^T gL), diag * 0.5 (bwd.cu:458-481)
            for c in range(r + 1):
                acc = zero
                for k in range(Cd):
                    acc = acc + L[k][r] * gL[k][c]
                S[r][c] = acc * 0.5 if r == c else acc
        U = [[zero for _ in range(Cd)] for _ in range(Cd)]
        for col in range(Cd):  # (L^T) U = S, bwd.cu:486-503
            for i in range(Cd - 1, -1, -1):
                s = S[i][col]
                for k in range(i + 1, Cd):
                    s = s - L[k][i] * U[k][col]
                U[i][col] = s / _guard(L[i][i])
        Gm = [[zero for _ in range(Cd)] for _ in range(Cd)]
        for col in range(Cd):  # the reference's "G = U L^-1" (solved with L, as written: bwd.cu:505-524)
            for i in range(Cd):
                s = U[col][i]
                for k in range(i):
                    s = s - L[i][k] * Gm[k][col]
                Gm[i][col] = s / _guard(L[i][i])
        for r in range(Cd):
            for c in range(Cd):
                gVfull[:, 3 + r, 3 + c] += 0.5 * (Gm[r][c] + Gm[c][r])  # bwd.cu:527-533
        gm[:, 3:] = -gx  # bwd.cu:538-539
        gb = gb + dL_dba * torch.where(betas < 4.0, torch.full_like(betas, 0.25), torch.zeros_like(betas))
        gVfull[:, 3:, 3:] += -(i22.transpose(1, 2) @ (Gi22 @ i22.transpose(1, 2)))  # bwd.cu:548-562
        return gm, gVfull, go, gb, None


def cond_mean_convariance_opacity(means, covars, opacities, betas, query):
    """K3/K4 (cuda/_wrapper.py:18-31,573-611).  means [N,D], covars [N,D,D], opacities [N,1], betas/query [N,Cd]
    -> [N,3], [N,3,3], [N,1].  No gradient to `query` (cuda/_wrapper.py:611)."""
    return _Conditioning.apply(means, covars, opacities, betas, query.detach())


def fully_fused_projection(means, covars6, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01, far_plane=1e10,
                           radius_clip=0.0, calc_compensations=False, extent=1.0):
    """K5 with the CUDA culling rules (fully_fused_projection_fwd.cu:43-177).  `extent` is the radius in sigmas:
    1.0 is the CUDA path (:147-150); 3.0 reproduces the reference's _torch_impl.py:364 for the golden fixtures.
    Returns radii [C,N] int32, means2d [C,N,2], depths [C,N], conics [C,N,3], compensations [C,N] or None;
    float outputs of culled entries are zeroed (the reference leaves them uninitialised)."""
    R, t = viewmats[:, :3, :3], viewmats[:, :3, 3]
    mc = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]  # utils.cuh:374-381
    xx, xy, xz, yy, yz, zz = covars6.unbind(-1)
    cov = torch.stack([xx, xy, xz, xy, yy, yz, xz, yz, zz], dim=-1).reshape(-1, 3, 3)
    cov_c = torch.einsum("cij,njk,clk->cnil", R, cov, R)  # utils.cuh:396-403
    x, y, z = mc.unbind(-1)
    fx, fy = Ks[:, 0, 0, None], Ks[:, 1, 1, None]
    cx, cy = Ks[:, 0, 2, None], Ks[:, 1, 2, None]
    tan_fovx, tan_fovy = 0.5 * width / fx, 0.5 * height / fy
    lim_x_pos, lim_x_neg = (width - cx) / fx + 0.3 * tan_fovx, cx / fx + 0.3 * tan_fovx
    lim_y_pos, lim_y_neg = (height - cy) / fy + 0.3 * tan_fovy, cy / fy + 0.3 * tan_fovy
    rz = 1.0 / z
    rz2 = rz * rz
    tx = z * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, x * rz))  # utils.cuh:276-277
    ty = z * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, y * rz))
    O = torch.zeros_like(z)
    J = torch.stack([fx * rz, O, -fx * tx * rz2, O, fy * rz, -fy * ty * rz2], dim=-1).reshape(*z.shape, 2, 3)
    cov2d = J @ cov_c @ J.transpose(-1, -2)
    means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)
    det_orig = cov2d[..., 0, 0] * cov2d[..., 1, 1] - cov2d[..., 0, 1] * cov2d[..., 1, 0]
    c00, c11, c01 = cov2d[..., 0, 0] + eps2d, cov2d[..., 1, 1] + eps2d, cov2d[..., 0, 1]
    det = c00 * c11 - c01 * cov2d[..., 1, 0]  # utils.cuh:458-466
    comp = torch.sqrt(torch.clamp(det_orig / det, min=0.0))
    safe_det = torch.where(det > 0, det, torch.ones_like(det))
    conics = torch.stack([c11 / safe_det, -c01 / safe_det, c00 / safe_det], dim=-1)  # utils.cuh:437-451
    b = 0.5 * (c00 + c11)
    v1 = b + torch.sqrt(torch.clamp(b * b - det, min=0.01))
    radius = torch.ceil(extent * torch.sqrt(v1)).detach()
    valid = ~((z < near_plane) | (z > far_plane)) & (det > 0) & ~(radius <= radius_clip)
    valid &= ~((means2d[..., 0] + radius <= 0) | (means2d[..., 0] - radius >= width) |
               (means2d[..., 1] + radius <= 0) | (means2d[..., 1] - radius >= height))
    radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)
    vf = valid.to(means.dtype)
    out = (radii, means2d * vf[..., None], z * vf, conics * vf[..., None],
           (comp * vf) if calc_compensations else None)
    return out


# ----------------------------------------------------------------------------------------------------------------
# tile lists (C oracle)
# ----------------------------------------------------------------------------------------------------------------
def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True):
    """isect_tiles.cu:16-285 -> tiles_per_gauss [C,N] i32, isect_ids [I] i64 (sorted), flatten_ids [I] i32."""
    L = lib()
    m2 = _np(means2d, np.float32)
    r = _np(radii, np.int32)
    d = _np(depths, np.float32)
    C, N = r.shape
    tpg = np.empty((C, N), np.int32)
    n = L.oracle_isect_count(c_int64(C * N), _p(m2), _p(r), c_int(tile_size), c_int(tile_width), c_int(tile_height),
                             _p(tpg))
    ids = np.empty((n,), np.int64)
    flat = np.empty((n,), np.int32)
    L.oracle_isect_emit(c_int64(C), c_int64(N), _p(m2), _p(r), _p(d), c_int(tile_size), c_int(tile_width),
                        c_int(tile_height), _p(tpg), _p(ids), _p(flat))
    if sort:
        tb = L.oracle_id_bits(ctypes.c_uint32(tile_width * tile_height))
        cb = L.oracle_id_bits(ctypes.c_uint32(C))
        L.oracle_sort_pairs(c_int64(n), _p(ids), _p(flat), c_int(32 + tb + cb))
    return torch.from_numpy(tpg), torch.from_numpy(ids), torch.from_numpy(flat)


def isect_offset_encode(isect_ids, n_cameras, tile_width, tile_height):
    ids = _np(isect_ids, np.int64)
    off = np.empty((n_cameras, tile_height, tile_width), np.int32)
    lib().oracle_offset_encode(c_int64(ids.size), _p(ids), c_int(n_cameras), c_int(tile_width), c_int(tile_height),
                               _p(off))
    return torch.from_numpy(off)


# ----------------------------------------------------------------------------------------------------------------
# compositing (C oracle)
# ----------------------------------------------------------------------------------------------------------------
def rasterize_fwd(means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size, offsets,
                  flatten_ids):
    """rasterize_to_pixels_fwd.cu:16-191 -> render_colors [C,H,W,CH], render_alphas [C,H,W,1], last_ids [C,H,W]."""
    m2, co, cl = _np(means2d, np.float32), _np(conics, np.float32), _np(colors, np.float32)
    op, be = _np(opacities, np.float32), _np(betas, np.float32)
    bg, mk = _np(backgrounds, np.float32), _np(masks, np.uint8)
    off, flat = _np(offsets, np.int32), _np(flatten_ids, np.int32)
    C, N, CH = cl.shape
    rc = np.zeros((C, height, width, CH), np.float32)
    ra = np.zeros((C, height, width, 1), np.float32)
    li = np.zeros((C, height, width), np.int32)
    lib().oracle_rasterize_fwd(c_int(C), c_int64(N), c_int64(flat.size), _p(m2), _p(co), _p(cl), _p(op), _p(be),
                               _p(bg), _p(mk), c_int(CH), c_int(width), c_int(height), c_int(tile_size), _p(off),
                               _p(flat), _p(rc), _p(ra), _p(li))
    return torch.from_numpy(rc), torch.from_numpy(ra), torch.from_numpy(li)


def rasterize_counts(means2d, conics, opacities, betas, width, height, tile_size, offsets, flatten_ids):
    """(evaluations reaching the sigma test, evaluations accepted) -- the E_test / E_acc of SURVEY.md 8(d)."""
    m2, co = _np(means2d, np.float32), _np(conics, np.float32)
    op, be = _np(opacities, np.float32), _np(betas, np.float32)
    off, flat = _np(offsets, np.int32), _np(flatten_ids, np.int32)
    a, b = c_int64(0), c_int64(0)
    lib().oracle_rasterize_counts(c_int(off.shape[0]), c_int64(flat.size), _p(m2), _p(co), _p(op), _p(be),
                                  c_int(width), c_int(height), c_int(tile_size), _p(off), _p(flat),
                                  ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def rasterize_bwd(means2d, conics, colors, opacities, betas, backgrounds, masks, width, height, tile_size, offsets,
                  flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas):
    """rasterize_to_pixels_bwd.cu:16-276 -> v_means2d, v_conics, v_colors, v_opacities, v_betas (float32)."""
    m2, co, cl = _np(means2d, np.float32), _np(conics, np.float32), _np(colors, np.float32)
    op, be = _np(opacities, np.float32), _np(betas, np.float32)
    bg, mk = _np(backgrounds, np.float32), _np(masks, np.uint8)
    off, flat = _np(offsets, np.int32), _np(flatten_ids, np.int32)
    ra, li = _np(render_alphas, np.float32), _np(last_ids, np.int32)
    vrc, vra = _np(v_render_colors, np.float32), _np(v_render_alphas, np.float32)
    C, N, CH = cl.shape
    g_m2, g_co = np.zeros((C, N, 2), np.float64), np.zeros((C, N, 3), np.float64)
    g_cl, g_op, g_be = np.zeros((C, N, CH), np.float64), np.zeros((C, N), np.float64), np.zeros((C, N), np.float64)
    lib().oracle_rasterize_bwd(c_int(C), c_int64(N), c_int64(flat.size), _p(m2), _p(co), _p(cl), _p(op), _p(be),
                               _p(bg), _p(mk), c_int(CH), c_int(width), c_int(height), c_int(tile_size), _p(off),
                               _p(flat), _p(ra), _p(li), _p(vrc), _p(vra), _p(g_m2), _p(g_co), _p(g_cl), _p(g_op),
                               _p(g_be))
    return tuple(torch.from_numpy(g.astype(np.float32)) for g in (g_m2, g_co, g_cl, g_op, g_be))


# ----------------------------------------------------------------------------------------------------------------
# the whole path: raw parameters + cameras -> image (+ parameter gradients)
# ----------------------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    """Glue: C compositing inside torch.autograd so that the torch per-primitive stages back-propagate from it."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, betas, backgrounds, width, height, tile_size, offsets,
                flatten_ids):
        rc, ra, li = rasterize_fwd(means2d, conics, colors, opacities, betas, backgrounds, None, width, height,
                                   tile_size, offsets, flatten_ids)
        ctx.save_for_backward(means2d, conics, colors, opacities, betas, backgrounds, offsets, flatten_ids, ra, li)
        ctx.dims = (width, height, tile_size)
        return rc, ra

    @staticmethod
    def backward(ctx, v_rc, v_ra):
        means2d, conics, colors, opacities, betas, backgrounds, offsets, flatten_ids, ra, li = ctx.saved_tensors
        width, height, tile_size = ctx.dims
        g = rasterize_bwd(means2d, conics, colors, opacities, betas, backgrounds, None, width, height, tile_size,
                          offsets, flatten_ids, ra, li, v_rc.contiguous(), v_ra.contiguous())
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[5]:
            v_bg = (v_rc * (1.0 - ra)).sum(dim=(1, 2))  # cuda/_wrapper.py:1029-1034
        return g[0], g[1], g[2], g[3], g[4], v_bg, None, None, None, None, None


def condition(params, cam_pos, timestamp):
    """BetaModel.render up to the rasterization() call (scene/beta_model.py:660-696) for one camera:
    params = (xyz, mean, rgb, opacity, beta, scale, l_triangle) raw tensors -> means3, covars3x3, opac [N], beta0 [N]."""
    xyz, mean, rgb, opacity, beta, scale, l_triangle = params
    D = scale.shape[1]
    s, o, b, m = activations(xyz, mean, opacity, beta, scale)
    rot = l_triangle_to_rotmat(l_triangle[:, :3])
    covar = rot_scale_l_triangle_to_covar(rot, s, l_triangle)
    view_dir = xyz - cam_pos[None, :]
    view_dir = view_dir / view_dir.norm(dim=-1, keepdim=True)
    query = view_dir if D == 6 else torch.cat([view_dir, torch.full((xyz.shape[0], 1), float(timestamp))], dim=-1)
    m3, v3, oc = cond_mean_convariance_opacity(m, covar, o, b[:, 1:], query)
    return m3, v3, oc[:, 0], b[:, 0]


def rasterization(means, covars, opacities, betas, colors, viewmats, Ks, width, height, near_plane=0.01,
                  far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16, backgrounds=None,
                  rasterize_mode="classic"):
    """The reference's rasterization() call sequence, RGB mode (submodules/gsplat/rendering.py:48-218)."""
    C = viewmats.shape[0]
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    covars6 = covars[..., tri[0], tri[1]]
    radii, means2d, depths, conics, comps = fully_fused_projection(
        means, covars6, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations=(rasterize_mode == "antialiased"))
    opac = opacities.repeat(C, 1)
    bet = betas.repeat(C, 1)
    if comps is not None:
        opac = opac * comps
    cols = colors.expand(C, -1, -1) if colors.dim() == 2 else colors
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tpg, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th)
    offsets = isect_offset_encode(isect_ids, C, tw, th)
    rc, ra = _Composite.apply(means2d, conics, cols.contiguous(), opac, bet, backgrounds, width, height, tile_size,
                              offsets, flatten_ids)
    meta = dict(radii=radii, means2d=means2d, depths=depths, conics=conics, opacities=opac, betas=bet,
                tiles_per_gauss=tpg, isect_ids=isect_ids, flatten_ids=flatten_ids, isect_offsets=offsets)
    return rc, ra, meta


def render(params, viewmat, K, cam_pos, timestamp, width, height, background=None, **kw):
    """BetaModel.render for one camera (scene/beta_model.py:660-722): returns render_colors [1,H,W,3],
    render_alphas [1,H,W,1], meta.  Differentiable w.r.t. the 7 raw parameter tensors."""
    m3, v3, oc, b0 = condition(params, cam_pos, timestamp)
    bg = None if background is None else background[None]
    return rasterization(m3, v3, oc, b0, params[2], viewmat[None], K[None], width, height, backgrounds=bg, **kw)
