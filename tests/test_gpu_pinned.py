"""GPU tests pinning the corners the first round left unpinned: antialiased mode (operator chain and
fused, forward and backward), tile masks, depth-channel VALUES of every depth mode through the operator chain, the
stale-frame and pair-overflow guards of the fused rasteriser, and bit-exact tile lists at the full BASELINE sizes."""
import warnings

import math

import pytest
import torch

pytestmark = pytest.mark.gpu
IMG_ATOL = 1e-4
TRI = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _inputs(N, seed, W, H, C=1):
    from test_gpu_forward_stages import _conditioned_inputs

    return _conditioned_inputs(N, seed, W, H, C)


# ---- antialiased ---------------------------------------------------------------------------------------------------
def test_antialiased_rasterization_forward_and_backward_match_reference_kernels():
    """rasterize_mode="antialiased" (rendering.py:95,104-105): opacities x compensation, gradients through both."""
    ref = _ref()
    C_ = ref.load()
    import ubs_b200
    from test_gpu_backward import _assert_grad_close

    N, W, H, C = 40000, 480, 360, 1
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 808, W, H, C)
    covars = covars * 0.05  # small footprints: the 0.3 px blur matters and compensations are well below 1
    bg = torch.tensor([[0.1, 0.2, 0.3]], device="cuda")
    leaves = [t.clone().requires_grad_(True) for t in (means, covars, opac, betas, colors)]
    rc, ra, meta = ubs_b200.rasterization(leaves[0], None, None, leaves[2], leaves[3], leaves[4], viewmats, Ks, W, H,
                                          backgrounds=bg, rasterize_mode="antialiased", covars=leaves[1])
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg,
                              calc_compensations=True)
    vis = R["radii"] > 0
    assert float(R["compensations"][vis].mean()) < 0.9, "compensations are trivial: not a test of the mode"
    assert torch.equal(meta["radii"], R["radii"]) and torch.equal(meta["flatten_ids"], R["flatten_ids"])
    torch.testing.assert_close(meta["opacities"][vis], R["opacities"][vis], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(rc, R["render_colors"], rtol=0, atol=IMG_ATOL)
    torch.testing.assert_close(ra, R["render_alphas"], rtol=0, atol=IMG_ATOL)

    g = torch.Generator(device="cuda").manual_seed(2)
    v_rc = torch.randn(C, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda", generator=g) / (H * W)
    torch.autograd.backward((rc, ra), (v_rc, v_ra))
    # the same graph through the reference kernels: K11, opacity = opac * comp, K6 with v_compensations
    g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(
        R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"],
        R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc, v_ra)
    v_comp = gop * opac[None]
    v_opac = (gop * R["compensations"]).sum(0)
    cov6 = covars[..., TRI[0], TRI[1]].contiguous()
    g_m, g_c6, _, _, _ = C_.fully_fused_projection_bwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3, False,
                                                       R["radii"], R["conics"], R["compensations"], g2d,
                                                       torch.zeros_like(R["depths"]), gcon, v_comp.contiguous(), False)
    g_cov = torch.zeros_like(covars)
    g_cov[:, TRI[0], TRI[1]] = g_c6
    for nm, mine, want in (("means", leaves[0].grad, g_m), ("covars", leaves[1].grad, g_cov),
                           ("opacities", leaves[2].grad, v_opac), ("betas", leaves[3].grad, gbe.sum(0)),
                           ("colors", leaves[4].grad, gcol.sum(0))):
        _assert_grad_close(nm, mine, want.reshape(mine.shape), rtol=2e-3)


def test_fused_antialiased_matches_reference_chain():
    ref = _ref()
    C_ = ref.load()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import fused, synth

    D, N, W, H = 6, 50000, 512, 384
    scene = synth.make_scene(N, D, seed=909).to("cuda")
    scene.scale[:, :3] -= 1.0  # smaller footprints (see above)
    cam = synth.make_cameras(1, W, H, seed=4, device="cuda")[0]
    bg = torch.tensor([0.4, 0.4, 0.1], device="cuda")
    m, v, o, b0 = ref.condition(scene, cam)
    R = ref.rasterization_fwd(m, v, o, b0, scene.rgb, cam.viewmat[None], cam.K[None], W, H, backgrounds=bg[None],
                              calc_compensations=True)
    vis = R["radii"] > 0
    assert float(R["compensations"][vis].mean()) < 0.95
    rec = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, antialiased=True)
    rc, ra = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None])
    assert (rz.radii == R["radii"]).float().mean().item() > 0.999
    torch.testing.assert_close(rc, R["render_colors"], rtol=0, atol=IMG_ATOL)
    torch.testing.assert_close(ra, R["render_alphas"], rtol=0, atol=IMG_ATOL)
    # backward: screen-space gradients through the reference kernels, then the conditioning chain by torch autograd
    g = torch.Generator(device="cuda").manual_seed(3)
    v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.zeros(1, H, W, 1, device="cuda")
    v_rec = rz.backward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None], v_rc, v_ra)
    g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(
        R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg[None], None, W, H, 16,
        R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc, v_ra)
    cov6 = v[..., TRI[0], TRI[1]].contiguous()
    g_m3, g_c6, _, _, _ = C_.fully_fused_projection_bwd(m, cov6, None, None, cam.viewmat[None], cam.K[None], W, H, 0.3,
                                                        False, R["radii"], R["conics"], R["compensations"], g2d,
                                                        torch.zeros_like(R["depths"]), gcon,
                                                        (gop * o[None]).contiguous(), False)
    # only the rgb and spatial-beta columns bypass the conditioning chain; they pin the antialiased compositing
    # backward, the conditioned-mean gradient pins the compensation VJP
    sl = fused.record_slices(D)
    _assert_grad_close("rgb", v_rec[:, sl["rgb"]], gcol[0], rtol=2e-3)
    beta0 = 4.0 * torch.exp(scene.beta[:, 0])
    _assert_grad_close("beta0", v_rec[:, sl["beta"]][:, 0], gbe[0] * beta0, rtol=2e-3)
    # xyz receives the conditioned-mean gradient unchanged (cond backward: g_mu1 = g_mean_c)
    _assert_grad_close("xyz", v_rec[:, sl["xyz"]], g_m3, rtol=3e-3)


# ---- tile masks ------------------------------------------------------------------------------------------------------
def test_tile_masks_forward_and_backward_match_reference_kernels():
    ref = _ref()
    C_ = ref.load()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import ops

    N, W, H, C = 20000, 320, 240, 2
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 606, W, H, C)
    bg = torch.rand(C, 3, device="cuda")
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    th, tw = R["isect_offsets"].shape[1:]
    masks = torch.rand(C, th, tw, device="cuda") < 0.6
    rc_r, ra_r, last_r = C_.rasterize_to_pixels_fwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"],
                                                    bg, masks, W, H, 16, R["isect_offsets"], R["flatten_ids"])
    rc, ra = ops.rasterize_to_pixels(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], W, H, 16,
                                     R["isect_offsets"], R["flatten_ids"], backgrounds=bg, masks=masks)
    pix = masks.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W]
    torch.testing.assert_close(rc[pix], rc_r[pix], rtol=0, atol=IMG_ATOL)
    torch.testing.assert_close(ra[pix], ra_r[pix], rtol=0, atol=IMG_ATOL)
    # masked-out tiles: background colour (rasterize_to_pixels_fwd.cu:73-79); alpha is not written by the reference
    assert torch.equal(rc[~pix], rc_r[~pix])
    v_rc = torch.randn(C, H, W, 3, device="cuda") / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda") / (H * W)
    r = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, masks, W, H,
                                   16, R["isect_offsets"], R["flatten_ids"], ra_r, last_r, v_rc, v_ra)
    mine = ops.rasterize_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg,
                             masks.contiguous().view(torch.uint8), W, H, 16, R["isect_offsets"], R["flatten_ids"], ra_r,
                             last_r, v_rc, v_ra)
    for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), mine, r):
        _assert_grad_close(name, a, b)


def test_gradient_row_kernel_with_tile_masks_matches_reference_backward():
    """ubs_rasterize_bwd_rows (the fused path's RGB backward) driven stand-alone with tile masks and two cameras, its rows
    decoded as include/ubs_b200.h documents, against the reference's rasterize_to_pixels_bwd on the same lists."""
    ref = _ref()
    C_ = ref.load()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import _lib
    from ubs_b200._lib import check, ptr

    N, W, H, C = 20000, 333, 250, 2
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 707, W, H, C)
    bg = torch.rand(C, 3, device="cuda")
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    th, tw = R["isect_offsets"].shape[1:]
    masks = torch.rand(C, th, tw, device="cuda") < 0.6
    rc_r, ra_r, last_r = C_.rasterize_to_pixels_fwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"],
                                                    bg, masks, W, H, 16, R["isect_offsets"], R["flatten_ids"])
    v_rc = torch.randn(C, H, W, 3, device="cuda") / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda") / (H * W)
    r = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, masks, W, H,
                                   16, R["isect_offsets"], R["flatten_ids"], ra_r, last_r, v_rc, v_ra)
    # the 48-byte splat rows the fused projection kernel would have written
    splats = torch.zeros(C, N, 12, device="cuda")
    splats[..., 0:2], splats[..., 2], splats[..., 3] = R["means2d"], R["opacities"], R["betas"]
    splats[..., 4:7], splats[..., 7], splats[..., 8:11] = R["conics"], R["depths"], R["colors"]
    rows = torch.zeros(C, N, 12, device="cuda")
    n = torch.tensor([R["flatten_ids"].numel()], dtype=torch.int64, device="cuda")
    m8 = masks.contiguous().view(torch.uint8)
    check(_lib.load().ubs_rasterize_bwd_rows(
        C, N, ptr(n), R["flatten_ids"].numel(), ptr(splats), ptr(bg), ptr(m8), W, H, 16, ptr(R["isect_offsets"]),
        ptr(R["flatten_ids"]), ptr(ra_r.contiguous()), ptr(last_r.contiguous()), ptr(v_rc), ptr(v_ra), ptr(rows), None,
        torch.cuda.current_stream().cuda_stream), "ubs_rasterize_bwd_rows")
    a, b, c = R["conics"].unbind(-1)
    mine = (torch.stack((2 * a * rows[..., 6] + 2 * b * rows[..., 7], 2 * b * rows[..., 6] + 2 * c * rows[..., 7]), -1),
            torch.stack((rows[..., 3], 2 * rows[..., 4], rows[..., 5]), -1), rows[..., 0:3], rows[..., 8],
            rows[..., 9] * math.log(2.0))
    for name, x, y in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), mine, r):
        _assert_grad_close(name, x, y.reshape(x.shape))
    assert (rows[..., 10:] == 0).all()


def _well_conditioned_normals(depth_img, c2w, Ks, noise=1e-4):
    """Pixels whose normal (direction of a cross product of depth differences) moves by < 5e-3 under the 1e-4
    depth noise the image tolerance allows: elsewhere the normal is the direction of a near-zero vector."""
    from ubs_b200 import rendering

    g = torch.Generator(device=depth_img.device).manual_seed(0)
    scale = max(float(depth_img.abs().max()), 1.0)
    n0 = rendering.depth_to_normal(depth_img, c2w, Ks)
    worst = torch.zeros_like(n0[..., 0])
    for _ in range(3):
        d = depth_img + noise * scale * (2 * torch.rand(depth_img.shape, device=depth_img.device, generator=g) - 1)
        worst = torch.maximum(worst, (rendering.depth_to_normal(d, c2w, Ks) - n0).abs().max(dim=-1).values)
    return (worst < 1e-2) & (n0.norm(dim=-1) > 0.5)


# ---- depth-channel values through the operator chain ----------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["RGB+D", "RGB+ED", "Depth", "EDepth", "Normal"])
def test_depth_channel_values_of_rasterization(mode):
    ref = _ref()
    import ubs_b200
    from ubs_b200 import rendering

    N, W, H, C = 50000, 480, 352, 1
    means, covars, opac, betas, colors, viewmats, Ks = _inputs(N, 4343, W, H, C)
    bg = torch.tensor([[0.2, 0.4, 0.6]], device="cuda")
    rc, ra, meta = ubs_b200.rasterization(means, None, None, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg,
                                          render_mode=mode, covars=covars)
    P = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    Rd = ref.rasterization_fwd(means, covars, opac, betas, P["depths"][0][:, None].repeat(1, 3), viewmats, Ks, W, H,
                               backgrounds=torch.zeros(1, 3, device="cuda"))
    acc = Rd["render_colors"][..., :1]  # accumulated depth over a zero background (rendering.py:131-142)
    tol = IMG_ATOL * max(float(acc.abs().max()), 1.0)
    if mode in ("Depth", "EDepth"):  # "EDepth" is NOT normalised: rendering.py:219 tests for "ED"
        torch.testing.assert_close(rc, acc, rtol=0, atol=tol)
    elif mode == "RGB+D":
        torch.testing.assert_close(rc[..., 3:], acc, rtol=0, atol=tol)
        torch.testing.assert_close(rc[..., :3], P["render_colors"], rtol=0, atol=IMG_ATOL)
    elif mode == "RGB+ED":
        ok = P["render_alphas"][..., 0] > 0.05
        want = acc / P["render_alphas"].clamp(min=1e-10)
        torch.testing.assert_close(rc[..., 3:][ok], want[ok], rtol=1e-3, atol=1e-3)
    else:
        c2w = torch.inverse(viewmats)
        want = (rendering.depth_to_normal(acc, c2w, Ks) + 1) / 2
        well = _well_conditioned_normals(acc, c2w, Ks)
        assert well.float().mean() > 0.05
        assert ((rc - want).abs().max(dim=-1).values[well] < 2e-2).float().mean() > 0.98


@pytest.mark.parametrize("channels", [4, 1])
def test_fused_depth_channel_backward_matches_reference_kernels(channels):
    """Gradients through the depth channel of the fused 4- / 1-channel modes: K11 of the reference with depth as a
    colour channel gives v_depths, which K6 consumes (cuda/_wrapper.py:898-925)."""
    ref = _ref()
    C_ = ref.load()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import fused, synth

    D, N, W, H = 6, 30000, 320, 256
    scene = synth.make_scene(N, D, seed=515).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=5, device="cuda")[0]
    m, v, o, b0 = ref.condition(scene, cam)
    P = ref.rasterization_fwd(m, v, o, b0, scene.rgb, cam.viewmat[None], cam.K[None], W, H)
    cols = torch.cat([P["colors"], P["depths"][..., None]], dim=-1) if channels == 4 else P["depths"][..., None]
    cols = cols.contiguous()
    bg = torch.zeros(1, channels, device="cuda")
    rc_r, ra_r, last_r = C_.rasterize_to_pixels_fwd(P["means2d"], P["conics"], cols, P["opacities"], P["betas"], bg,
                                                    None, W, H, 16, P["isect_offsets"], P["flatten_ids"])
    rec = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H)
    rc, ra = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg, channels=channels)
    scale = max(float(rc_r.abs().max()), 1.0)
    torch.testing.assert_close(rc, rc_r, rtol=0, atol=IMG_ATOL * scale)
    g = torch.Generator(device="cuda").manual_seed(8)
    v_rc = torch.randn(1, H, W, channels, device="cuda", generator=g) / (H * W)
    v_ra = torch.zeros(1, H, W, 1, device="cuda")
    v_rec = rz.backward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg, v_rc, v_ra)
    g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(P["means2d"], P["conics"], cols, P["opacities"], P["betas"],
                                                           bg, None, W, H, 16, P["isect_offsets"], P["flatten_ids"],
                                                           ra_r, last_r, v_rc, v_ra)
    v_depths = gcol[..., -1].contiguous()
    cov6 = v[..., TRI[0], TRI[1]].contiguous()
    g_m3, _, _, _, _ = C_.fully_fused_projection_bwd(m, cov6, None, None, cam.viewmat[None], cam.K[None], W, H, 0.3,
                                                     False, P["radii"], P["conics"], None, g2d, v_depths, gcon, None,
                                                     False)
    sl = fused.record_slices(D)
    _assert_grad_close("xyz (incl. v_depths)", v_rec[:, sl["xyz"]], g_m3, rtol=3e-3)
    if channels == 4:
        _assert_grad_close("rgb", v_rec[:, sl["rgb"]], gcol[0, :, :3], rtol=2e-3)
    # and the depth gradient really matters in this test
    g_m3_nod, _, _, _, _ = C_.fully_fused_projection_bwd(m, cov6, None, None, cam.viewmat[None], cam.K[None], W, H, 0.3,
                                                         False, P["radii"], P["conics"], None, g2d,
                                                         torch.zeros_like(v_depths), gcon, None, False)
    assert float((g_m3 - g_m3_nod).abs().max()) > 2e-3 * float(g_m3.abs().max())


# ---- guards of the fused rasteriser ------------------------------------------------------------------------------------
def test_stale_frame_backward_raises():
    from ubs_b200 import UbsError, fused, synth

    D, N, W, H = 6, 5000, 128, 96
    scene = synth.make_scene(N, D, seed=1).to("cuda")
    cams = synth.make_cameras(2, W, H, seed=2, device="cuda")
    rec = fused.pack_records(D, *scene.tensors()).requires_grad_(True)
    rz = fused.FusedRasterizer(D, N, W, H)
    frames = [fused.render(rec, rz, c.viewmat[None], c.K[None], c.cam_pos[None]) for c in cams]
    with pytest.raises(UbsError, match="stale frame"):
        frames[0][0].sum().backward()
    frames[1][0].sum().backward()  # the most recent frame is fine
    assert rec.grad is not None and float(rec.grad.abs().sum()) > 0


def test_pair_overflow_is_never_silent_and_never_trains():
    from ubs_b200 import fused, synth, training

    D, N, W, H = 6, 20000, 256, 192
    scene = synth.make_scene(N, D, seed=3).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=4, device="cuda")[0]
    rec = fused.pack_records(D, *scene.tensors())
    args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, torch.zeros(1, 3, device="cuda"))
    full = fused.FusedRasterizer(D, N, W, H)
    gt = full.forward(rec, *args)[0].clone().permute(0, 3, 1, 2).contiguous() * 0.5
    n_pairs = full.last_pair_count()
    assert n_pairs > 5000
    rz = fused.FusedRasterizer(D, N, W, H, capacity=n_pairs // 3)
    adam = training.PackedAdam(D, N)
    step = training.TrainStep(rz, adam)
    before = rec.clone()
    step.step(rec, *args, gt)  # truncated frame: the fused Adam update must not be applied
    assert rz.overflowed() and int(rz.status[1]) == 1
    assert torch.equal(rec, before), "a truncated frame updated the parameters"
    assert float(adam.exp_avg.abs().sum()) == 0.0
    torch.cuda.synchronize()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        step.step(rec, *args, gt)  # the host has seen the count by now: warns, grows, and this frame is complete
    assert any("truncated" in str(x.message) for x in w)
    assert rz.capacity >= n_pairs and not rz.overflowed() and rz.truncated_frames == 1
    assert not torch.equal(rec, before)
    # non-fused backward of a truncated frame: zero gradient (the view is dropped), not a wrong one
    rz2 = fused.FusedRasterizer(D, N, W, H, capacity=n_pairs // 3, on_overflow="raise")
    rc, ra = rz2.forward(before, *args)
    v = rz2.backward(before, *args, torch.ones_like(rc), torch.zeros_like(ra))
    assert float(v.abs().sum()) == 0.0
    torch.cuda.synchronize()
    from ubs_b200 import UbsError

    with pytest.raises(UbsError, match="truncated"):
        rz2.forward(before, *args)


# ---- bit-exact tile lists at the full sizes --------------------------------------------------------------------------
@pytest.mark.parametrize("name,C", [("cfg3", 1), ("cfg5", 4)])
def test_full_size_tile_lists_are_bit_exact_on_reference_projection_outputs(name, C):
    """Stage-isolated at BASELINE sizes (3M primitives, 1920x1080; cfg5 with C=4 cameras in one call): the reference's
    own projection outputs into both sort routes -> every integer output equals the reference's K7-K9."""
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops, synth

    scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=64 if name == "cfg5" else C)
    W, H = cfg["width"], cfg["height"]
    pick = cams[:C] if name != "cfg5" else [cams[k] for k in (1, 18, 35, 60)]
    m, v, o, b0 = ref.condition(scene, pick[0])  # any consistent projection input will do
    cov6 = v[..., TRI[0], TRI[1]].contiguous()
    V = torch.stack([c.viewmat for c in pick]).contiguous()
    K = torch.stack([c.K for c in pick]).contiguous()
    radii, means2d, depths, conics, _ = C_.fully_fused_projection_fwd(m, cov6, None, None, V, K, W, H, 0.3, 0.01, 1e10,
                                                                      0.0, False, False)
    tw, th = (W + 15) // 16, (H + 15) // 16
    tpg_r, ids_r, fl_r = C_.isect_tiles(means2d, radii, depths, None, None, C, 16, tw, th, True, True)
    off_r = C_.isect_offset_encode(ids_r, C, tw, th)
    assert ids_r.numel() > 4 * 1000 * 1000 * (1 if C == 1 else 3)
    for method in ("bin", "onesweep"):
        tpg, ids, fl, off = ops.isect_tiles(means2d, radii, depths, 16, tw, th, n_cameras=C, return_offsets=True,
                                            method=method)
        for nm, a, b in (("tiles_per_gauss", tpg, tpg_r), ("isect_ids", ids, ids_r), ("flatten_ids", fl, fl_r),
                         ("isect_offsets", off, off_r)):
            assert a.shape == b.shape, (method, nm, a.shape, b.shape)
            assert torch.equal(a, b), "%s/%s differs at %d of %d entries" % (method, nm, int((a != b).sum()), a.numel())
        del tpg, ids, fl, off
