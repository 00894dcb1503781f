"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the train-step pieces around the render.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product
(universal-beta-splatting_b200/) never does.

  * l1_loss / ssim            utils/loss_utils.py:18-19,26-85 (the training loop calls the third-party `fused_ssim`
                              CUDA package, train.py:120, pinned rahul-goel/fused-ssim@1272e21 in setup.py:26 and
                              absent from /root/reference; its published algorithm is this same windowed SSIM with
                              zero "same" padding).  Pinned by tests/golden/loss_ssim.npz = outputs of the
                              reference's own utils/loss_utils.py imported in the build container
                              (tests/golden/make_golden_loss.py).
  * photometric_loss          train.py:118-121
  * regularisers              train.py:122-124 (note `get_scale[:3]`: the first three primitives)
  * make_adam                 scene/beta_model.py:239-268 -- torch.optim.Adam itself is the oracle of the fused Adam
  * update_params / relocate / add_new   scene/beta_model.py:548-657 with the sampled indices given.  Pinned by
                              tests/golden/relocate_D{6,7}.npz = inputs, drawn indices and outputs (parameters and Adam
                              moments) of the reference's own relocate_gs / add_new_gs, whose source text
                              tests/golden/make_golden_relocate.py compiles unmodified into a stub class (the module
                              itself cannot be imported here: plyfile / fused_ssim / the CUDA extension are missing)
  * sgld_noise                train.py:156-163 on the K1 / K2 restatements of oracle/ubs_oracle.py.  Pinned by
                              tests/golden/sgld_D{6,7}.npz = those four statements of the reference's train.py,
                              executed unmodified by tests/golden/make_golden_sgld.py
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def _window(window_size, channel, dtype):
    g = torch.tensor([exp(-((x - window_size // 2) ** 2) / float(2 * 1.5 ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    w2d = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2d.expand(channel, 1, window_size, window_size).contiguous().to(dtype)


def ssim(img1, img2, window_size=11):
    """[B,ch,H,W] (or [ch,H,W]) -> mean SSIM (utils/loss_utils.py:45-85, size_average=True)."""
    channel = img1.size(-3)
    window = _window(window_size, channel, img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def photometric_loss(image, gt_image, lambda_dssim=0.2):
    """train.py:118-121 for [ch,H,W] or [B,ch,H,W] images."""
    if image.dim() == 3:
        image, gt_image = image.unsqueeze(0), gt_image.unsqueeze(0)
    return (1.0 - lambda_dssim) * l1_loss(image, gt_image) + lambda_dssim * (1.0 - ssim(image, gt_image))


def regularisers(raw_opacity, raw_scale, opacity_reg=0.01, scale_reg=0.01):
    """train.py:122-124: opacity_reg * |sigmoid(o)|.mean() + scale_reg * |softplus(s)[:3]|.mean()."""
    return (opacity_reg * torch.abs(torch.sigmoid(raw_opacity)).mean()
            + scale_reg * torch.abs(F.softplus(raw_scale)[:3]).mean())


GROUPS = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")


def make_adam(params, lr):
    """params: the seven leaf tensors in BetaModel order; lr: dict group -> learning rate."""
    groups = [{"params": [p], "lr": lr[n], "name": n} for n, p in zip(GROUPS, params)]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


def update_params(params, idxs, ratio):
    """_update_params (scene/beta_model.py:548-565). params in BetaModel order; returns the 7 new tensors."""
    xyz, mean, rgb, opacity, beta, scale, l_triangle = params
    new_opacity = 1.0 - torch.pow(1.0 - torch.sigmoid(opacity[idxs, 0]), 1.0 / (ratio + 1))
    new_opacity = torch.clamp(new_opacity.unsqueeze(-1), max=1.0 - torch.finfo(torch.float32).eps, min=0.005)
    new_opacity = torch.log(new_opacity / (1 - new_opacity))
    return xyz[idxs], mean[idxs], rgb[idxs], new_opacity, beta[idxs], scale[idxs], l_triangle[idxs]


def relocate(params, moments, dead_indices, reinit_idx):
    """relocate_gs (scene/beta_model.py:575-620) with the sampled `reinit_idx` given; in place.
    moments: list of (exp_avg, exp_avg_sq) per tensor, or None."""
    ratio = torch.bincount(reinit_idx)[reinit_idx]
    new = update_params(params, reinit_idx, ratio)
    for p, n in zip(params, new):
        p.index_copy_(0, dead_indices, n)
    params[3].index_copy_(0, reinit_idx, params[3].index_select(0, dead_indices))
    if moments is not None:
        for m, v in moments:
            m[reinit_idx] = 0
            v[reinit_idx] = 0


def add_new(params, moments, add_idx):
    """add_new_gs (scene/beta_model.py:622-657) with the sampled `add_idx` given; returns (params, moments) grown."""
    ratio = torch.bincount(add_idx)[add_idx]
    new = update_params(params, add_idx, ratio)
    params[3][add_idx] = new[3]
    out = [torch.cat((p, n)) for p, n in zip(params, new)]
    out_m = None
    if moments is not None:
        out_m = []
        for (m, v), n in zip(moments, new):
            m2, v2 = torch.cat((m, torch.zeros_like(n))), torch.cat((v, torch.zeros_like(n)))
            m2[add_idx] = 0
            v2[add_idx] = 0
            out_m.append((m2, v2))
    return out, out_m


def sgld_noise(params, noise, noise_lr, xyz_lr):
    """train.py:156-163 with the N(0,1) draw given: returns the new xyz.  get_xyz_covariance
    (scene/beta_model.py:143-152) is the spatial block of rot_scale_l_triangle_to_covar on get_rotation / get_scale /
    get_l_triangle -- the K1 / K2 restatements of oracle/ubs_oracle.py, which tests/test_oracle_golden.py pins against
    the reference's _torch_impl fixtures (spatial_block included)."""
    from oracle import ubs_oracle as O

    xyz, mean, rgb, opacity, beta, scale, l_triangle = params
    N = xyz.shape[0]
    cov = O.rot_scale_l_triangle_to_covar(O.l_triangle_to_rotmat(l_triangle[:, :3]), F.softplus(scale), l_triangle,
                                          spatial_block=True)
    n = torch.randn_like(xyz) if noise is None else noise
    n = n * torch.pow(1 - torch.sigmoid(opacity.reshape(N, 1)), 100) * noise_lr * xyz_lr
    return xyz + torch.bmm(cov, n.unsqueeze(-1)).squeeze(-1)
