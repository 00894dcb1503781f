"""TEST INFRASTRUCTURE ONLY -- loader for the reference's own CUDA kernels (oracle/_ref/ubs_ref_cuda.so).

The .so is built by oracle/build_ref.py from the sources under /root/reference (never copied here) and exposes
the reference's raw `_C` functions (csrc/bindings.h:34-273).  Only tests/, __graft_entry__.smoke() and bench.py's
reference arm may import this module.  The small wrappers below restate the *call sequence* of the reference's
Python glue (rendering.py:17-232, scene/beta_model.py:660-711) on top of those raw functions.
"""
import importlib.util
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "ubs_ref_cuda.so")
_C = None


def available():
    return os.path.exists(SO_PATH)


def load():
    global _C
    if _C is None:
        if not available():
            raise RuntimeError("reference oracle %s has not been built (python oracle/build_ref.py)" % SO_PATH)
        spec = importlib.util.spec_from_file_location("ubs_ref_cuda", SO_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _C = mod
    return _C


def tril_rest(D, device):
    """rest_i / rest_j exactly as the reference caller builds them (scene/beta_model.py:69-73)."""
    ti, tj = torch.tril_indices(D, D, offset=-1)
    m = (ti >= 3) | (tj >= 3)
    return ti[m].to(torch.int32).to(device), tj[m].to(torch.int32).to(device)


def activations(scene):
    """scene/beta_model.py:36-52,103-121."""
    scale = torch.nn.functional.softplus(scene.scale)
    opacity = torch.sigmoid(scene.opacity)
    beta = 4.0 * torch.exp(scene.beta)
    mean = torch.cat([scene.xyz, scene.mean], dim=-1)
    return scale, opacity, beta, mean


def query_for(scene, cam):
    """scene/beta_model.py:675-690."""
    view_dir = scene.xyz - cam.cam_pos.unsqueeze(0)
    view_dir = view_dir / view_dir.norm(dim=-1, keepdim=True)
    if scene.D == 6:
        return view_dir
    ts = torch.full((view_dir.shape[0], 1), cam.timestamp, device=view_dir.device, dtype=view_dir.dtype)
    return torch.cat([view_dir, ts], dim=-1)


@torch.no_grad()
def condition(scene, cam):
    """K1 -> K2 -> K3 through the reference kernels: returns means[N,3], covars[N,3,3], opac[N], beta0[N]."""
    C = load()
    scale, opacity, beta, mean = activations(scene)
    rot = C.l_triangle_to_rotmat_fwd(scene.l_triangle[:, :3].contiguous())
    ri, rj = tril_rest(scene.D, scene.xyz.device)
    covar = C.rot_scale_l_triangle_to_covar_fwd(rot, scale.contiguous(), scene.l_triangle.contiguous(), ri, rj, False)
    q = query_for(scene, cam).contiguous()
    m, v, o = C.cond_mean_convariance_opacity_fwd(mean.contiguous(), covar, opacity.contiguous(),
                                                  beta[:, 1:].contiguous(), q)
    return m, v, o.squeeze(-1), beta[:, 0].contiguous()


@torch.no_grad()
def rasterization_fwd(means, covars3x3, opacities, betas, colors, viewmats, Ks, width, height, backgrounds=None,
                      near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16,
                      calc_compensations=False):
    """Forward call sequence of the reference rasterization() (rendering.py:48-218), RGB mode."""
    C_ = load()
    Cn, N = viewmats.shape[0], means.shape[0]
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    covars6 = covars3x3[..., tri[0], tri[1]].contiguous()
    radii, means2d, depths, conics, comps = C_.fully_fused_projection_fwd(
        means.contiguous(), covars6, None, None, viewmats.contiguous(), Ks.contiguous(), width, height, eps2d,
        near_plane, far_plane, radius_clip, calc_compensations, False)
    opac = opacities.repeat(Cn, 1)
    bet = betas.repeat(Cn, 1)
    if calc_compensations:
        opac = opac * comps
    cols = colors.expand(Cn, -1, -1).contiguous() if colors.dim() == 2 else colors.contiguous()
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tiles_per_gauss, isect_ids, flatten_ids = C_.isect_tiles(means2d, radii, depths, None, None, Cn, tile_size, tw, th,
                                                             True, True)
    offsets = C_.isect_offset_encode(isect_ids, Cn, tw, th)
    rc, ra, last_ids = C_.rasterize_to_pixels_fwd(means2d, conics, cols, opac.contiguous(), bet.contiguous(),
                                                  backgrounds, None, width, height, tile_size, offsets, flatten_ids)
    return dict(radii=radii, means2d=means2d, depths=depths, conics=conics, compensations=comps, opacities=opac,
                betas=bet, colors=cols, tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids, flatten_ids=flatten_ids,
                isect_offsets=offsets, render_colors=rc, render_alphas=ra, last_ids=last_ids)


def chain_grads(scene, cam, bg, v_rc, v_ra, keep=None):
    """Gradients of the 7 raw parameter tensors through the reference kernels + torch glue, restating the autograd
    graph of scene/beta_model.py:660-711 with explicit VJP calls (cuda/_wrapper.py:573-684,807-1050)."""
    C_ = load()
    D = scene.D
    W, H = cam.width, cam.height
    raw = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
    xyz, mean, rgb, opacity, beta, scale, ltri = raw
    s_act = torch.nn.functional.softplus(scale)
    o_act = torch.sigmoid(opacity)
    b_act = 4.0 * torch.exp(beta)
    mu = torch.cat([xyz, mean], dim=-1)
    ri, rj = tril_rest(D, "cuda")
    with torch.no_grad():
        l3 = ltri[:, :3].contiguous()
        rot = C_.l_triangle_to_rotmat_fwd(l3)
        covar = C_.rot_scale_l_triangle_to_covar_fwd(rot, s_act.contiguous(), ltri.contiguous(), ri, rj, False)
        q = query_for(scene, cam).contiguous()
        bc = b_act[:, 1:].contiguous()
        m3, v3, o3 = C_.cond_mean_convariance_opacity_fwd(mu.contiguous(), covar, o_act.contiguous(), bc, q)
        R = rasterization_fwd(m3, v3, o3.squeeze(-1), b_act[:, 0].contiguous(), rgb, cam.viewmat[None], cam.K[None],
                                  W, H, backgrounds=bg[None])
        g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(
            R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg[None], None, W, H, 16,
            R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc, v_ra)
        tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        cov6 = v3[..., tri[0], tri[1]].contiguous()
        g_m3, g_cov6, _, _, g_view = C_.fully_fused_projection_bwd(
            m3, cov6, None, None, cam.viewmat[None], cam.K[None], W, H, 0.3, False, R["radii"], R["conics"], None, g2d,
            torch.zeros_like(R["depths"]), gcon, None, keep is not None)  # + v_viewmats for the callers that keep them
        g_v3 = torch.zeros_like(v3)
        g_v3[:, tri[0], tri[1]] = g_cov6
        g_mu, g_covar, g_o, g_bc = C_.cond_mean_convariance_opacity_bwd(
            mu.contiguous(), covar, o_act.contiguous(), bc, q, g_m3, g_v3.contiguous(), gop[0][:, None].contiguous())
        g_rot, g_s, g_lt = C_.rot_scale_l_triangle_to_covar_bwd(rot, s_act.contiguous(), ltri.contiguous(), ri, rj,
                                                               False, g_covar)
        g_l3 = C_.l_triangle_to_rotmat_bwd(l3, g_rot)
        g_lt = g_lt.clone()
        g_lt[:, :3] += g_l3
        g_b = torch.cat([gbe[0][:, None], g_bc], dim=-1)
        if keep is not None:  # intermediate results for the golden fixtures
            keep.update(cond_means=m3, cond_covars=v3, cond_opac=o3, v_means2d=g2d, v_conics=gcon, v_colors=gcol,
                        v_opacities=gop, v_betas=gbe, v_cond_means=g_m3, v_cond_cov6=g_cov6, v_mu=g_mu,
                        v_covar=g_covar, v_opac_act=g_o, v_beta_cond=g_bc, v_scale_act=g_s, v_rot=g_rot,
                        v_viewmats=g_view)
    # torch glue backward (activations, cat)
    torch.autograd.backward([s_act, o_act, b_act, mu], [g_s, g_o, g_b, g_mu])
    return [xyz.grad, mean.grad, gcol[0], opacity.grad, beta.grad, scale.grad, g_lt], R
