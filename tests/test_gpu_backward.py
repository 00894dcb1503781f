"""GPU parity of the backward stages (K11 compositing bwd, K6 projection bwd) and of autograd through
rasterization() against the reference's own CUDA kernels.  Bar: 1e-3 relative (north star)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
GRAD_RTOL = 1e-3


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _inputs(N, seed, W, H, C=1):
    from test_gpu_forward_stages import _conditioned_inputs

    return _conditioned_inputs(N, seed, W, H, C)


def _assert_grad_close(name, mine, theirs, rtol=GRAD_RTOL):
    """Gradients are sums of many float terms accumulated in a different order (atomics): compare against the
    tensor's scale, and require the bulk of the entries to agree to rtol individually."""
    scale = theirs.abs().max().clamp_min(1e-20)
    err = (mine - theirs).abs().max() / scale
    assert err < rtol, "%s: max abs err / max |ref| = %.3e" % (name, err.item())
    big = theirs.abs() > 1e-3 * scale
    if big.any():
        rel = ((mine - theirs).abs() / theirs.abs().clamp_min(1e-30))[big]
        frac_bad = (rel > 10 * rtol).float().mean().item()
        assert frac_bad < 1e-3, "%s: %.3f%% of significant entries off by > %g" % (name, 100 * frac_bad, 10 * rtol)


@pytest.mark.parametrize("N,W,H,C,bg", [(20000, 320, 240, 1, True), (60000, 640, 480, 1, False),
                                         (8000, 200, 136, 3, True)])
def test_rasterize_bwd_matches_reference(N, W, H, C, bg):
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 555 + N, W, H, C)
    opac = opac * 0.9 + 0.1
    backgrounds = torch.rand(C, 3, device="cuda") if bg else None
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=backgrounds)
    g = torch.Generator(device="cuda").manual_seed(1)
    v_rc = torch.randn(C, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda", generator=g) / (H * W)
    r = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], backgrounds,
                                   None, W, H, 16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"],
                                   R["last_ids"], v_rc, v_ra)
    m = ops.rasterize_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], backgrounds, None, W, H,
                          16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc, v_ra)
    for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), m, r):
        assert a.shape == b.shape
        _assert_grad_close(name, a, b)


@pytest.mark.parametrize("ch", [1, 4, 8])
def test_rasterize_bwd_other_channel_counts(ch):
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops

    N, W, H, C = 15000, 256, 192, 1
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 31 + ch, W, H, C)
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H)
    cols = torch.rand(C, N, ch, device="cuda")
    bgs = torch.rand(C, ch, device="cuda")
    rc_r, ra_r, last_r = C_.rasterize_to_pixels_fwd(R["means2d"], R["conics"], cols, R["opacities"], R["betas"], bgs,
                                                    None, W, H, 16, R["isect_offsets"], R["flatten_ids"])
    rc, ra, last = ops.rasterize_fwd(R["means2d"], R["conics"], cols, R["opacities"], R["betas"], bgs, None, W, H, 16,
                                     R["isect_offsets"], R["flatten_ids"])
    torch.testing.assert_close(rc, rc_r, rtol=0, atol=1e-4)
    v_rc = torch.randn(C, H, W, ch, device="cuda") / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda") / (H * W)
    r = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], cols, R["opacities"], R["betas"], bgs, None, W, H, 16,
                                   R["isect_offsets"], R["flatten_ids"], ra_r, last_r, v_rc, v_ra)
    m = ops.rasterize_bwd(R["means2d"], R["conics"], cols, R["opacities"], R["betas"], bgs, None, W, H, 16,
                          R["isect_offsets"], R["flatten_ids"], ra_r, last_r, v_rc, v_ra)
    for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), m, r):
        _assert_grad_close(name + "[ch=%d]" % ch, a, b)


@pytest.mark.parametrize("N,W,H,C,comp", [(30000, 400, 300, 1, False), (9000, 200, 136, 3, True)])
def test_projection_bwd_matches_reference(N, W, H, C, comp):
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 808 + N, W, H, C)
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    cov6 = covars[..., tri[0], tri[1]].contiguous()
    radii, m2d, depth, conic, comps = C_.fully_fused_projection_fwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3,
                                                                    0.01, 1e10, 0.0, comp, False)
    v_m2d = torch.randn_like(m2d)
    v_depth = torch.randn_like(depth)
    v_conic = torch.randn_like(conic)
    v_comp = torch.randn_like(depth) if comp else None
    r_means, r_cov, _, _, r_view = C_.fully_fused_projection_bwd(
        means, cov6, None, None, viewmats, Ks, W, H, 0.3, False, radii, conic, comps if comp else None, v_m2d,
        v_depth, v_conic, v_comp, True)
    v_means, v_cov, v_view = ops.projection_bwd(means, cov6, viewmats, Ks, W, H, 0.3, radii, conic,
                                                comps if comp else None, v_m2d, v_depth, v_conic, v_comp, True)
    _assert_grad_close("v_means", v_means, r_means)
    _assert_grad_close("v_covars", v_cov, r_cov)
    _assert_grad_close("v_viewmats", v_view[:, :3, :], r_view[:, :3, :], rtol=5e-3)
    assert (v_means[(radii <= 0).all(0)] == 0).all()


def _assert_grad_close_bulk(name, mine, theirs, rtol, max_outlier_frac=1e-4):
    """End-to-end variant: the forward intermediates of the two chains differ in the last bit (conics ~1e-7), which
    can move a pixel across the support boundary sigma = 1 of a primitive.  With beta < 1 the Beta kernel's
    derivative (1 - sigma)^(beta - 1) is unbounded there, so that one pixel changes that one primitive's gradient
    by tens of percent in EITHER implementation (measured: stage-by-stage on identical inputs the two agree to
    5e-7 of scale, scratch/flaky2.py).  Hence: all but a 1e-4 fraction of the entries within rtol of the scale."""
    scale = theirs.abs().max().clamp_min(1e-20)
    bad = ((mine - theirs).abs() > rtol * scale).float().mean().item()
    assert bad <= max_outlier_frac, "%s: %.4f%% of entries off by > %g of scale" % (name, 100 * bad, rtol)
    assert ((mine - theirs).abs().max() / scale) < 0.05, "%s: an entry is off by more than 5%% of scale" % name


def test_rasterization_autograd_end_to_end():
    ref = _ref()
    C_ = ref.load()
    import ubs_b200

    torch.manual_seed(20251003)

    N, W, H, C = 40000, 480, 360, 1
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 2024, W, H, C)
    bg = torch.tensor([[1.0, 1.0, 1.0]], device="cuda")
    leaves = [t.detach().clone().requires_grad_(True) for t in (means, covars, opac, betas, colors)]
    rc, ra, meta = ubs_b200.rasterization(leaves[0], None, None, leaves[2], leaves[3], leaves[4], viewmats, Ks, W, H,
                                          backgrounds=bg, covars=leaves[1])
    v_rc = torch.randn_like(rc) / (H * W)
    v_ra = torch.randn_like(ra) / (H * W)
    torch.autograd.backward((rc, ra), (v_rc, v_ra))

    # the reference: same chain through its raw kernels
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(
        R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"],
        R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc.contiguous(), v_ra.contiguous())
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    cov6 = covars[..., tri[0], tri[1]].contiguous()
    r_means, r_cov6, _, _, _ = C_.fully_fused_projection_bwd(
        means, cov6, None, None, viewmats, Ks, W, H, 0.3, False, R["radii"], R["conics"], None, g2d,
        torch.zeros_like(R["depths"]), gcon, None, False)
    r_cov = torch.zeros(N, 3, 3, device="cuda")
    r_cov[:, tri[0], tri[1]] = r_cov6  # index-backward of the 3x3 -> 6 gather: upper triangle only
    # two chained stages: 1e-3 of the tensor scale for all but isolated boundary-pixel outliers (see
    # _assert_grad_close_bulk); the stage-by-stage tests above hold the strict 1e-3 bar on identical inputs
    _assert_grad_close_bulk("means", leaves[0].grad, r_means, rtol=1e-3)
    _assert_grad_close_bulk("covars", leaves[1].grad, r_cov, rtol=1e-3)
    _assert_grad_close("opacities", leaves[2].grad, gop.sum(0))
    _assert_grad_close("betas", leaves[3].grad, gbe.sum(0))
    _assert_grad_close("colors", leaves[4].grad, gcol.sum(0))
