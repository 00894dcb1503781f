"""GPU tests of the zero-edit drop-in route: the reference caller's statements (tests/ref_caller.py;
AST-identical to scene/beta_model.py:103-159,660-722 -- tests/test_dropin_host.py) run against the gsplat shim must
reach the fused kernels and reproduce the reference's own CUDA kernels (oracle/_ref) driven by the same statements.

Because the route feeds the fused kernels the very tensors the reference's operators would receive (torch's softplus /
sigmoid / exp, the caller's own query), radii, depths and the tile lists are BIT-exact here, not just close."""
import pytest
import torch

pytestmark = pytest.mark.gpu
IMG_ATOL = 1e-4


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _caller(scene, bg, grad):
    import ref_caller

    return ref_caller, ref_caller.BetaModelCaller(scene, bg, requires_grad=grad)


def _ref_cam(cam, vc):
    from ubs_b200 import synth

    return synth.Camera(cam.viewmat, vc.K(), cam.cam_pos, cam.width, cam.height, cam.timestamp)


class _NoSync:
    """Any host synchronisation inside raises (torch.cuda.set_sync_debug_mode)."""

    def __enter__(self):
        torch.cuda.set_sync_debug_mode("error")

    def __exit__(self, *exc):
        torch.cuda.set_sync_debug_mode("default")
        return False


@pytest.mark.parametrize("D,N,W,H", [(6, 60000, 640, 480), (7, 40000, 507, 380)])
def test_reference_caller_statements_take_the_fused_route(D, N, W, H, monkeypatch):
    ref = _ref()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import dropin, synth

    scene = synth.make_scene(N, D, seed=77 + D).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=3, timestamps=[0.4], device="cuda")[0]
    bg = torch.tensor([0.2, 0.5, 0.1], device="cuda")
    rc_mod, model = _caller(scene, bg, grad=True)
    vc = rc_mod.ViewpointCamera(cam)

    # the shim's rasterization() itself must not synchronise the host (the caller's own `[mask]` gathers do)
    inner = rc_mod.rasterization

    def guarded(*a, **kw):
        with _NoSync():
            return inner(*a, **kw)

    monkeypatch.setattr(rc_mod, "rasterization", guarded)
    inner_bwd = dropin._DropinRender.backward

    def guarded_bwd(ctx, *grads):  # likewise the route's autograd node (the caller's `[mask]` backward does synchronise)
        with _NoSync():
            return inner_bwd(ctx, *grads)

    monkeypatch.setattr(dropin._DropinRender, "backward", staticmethod(guarded_bwd))
    before = dropin.stats()
    out = model.render(vc)
    after = dropin.stats()
    assert after["fused"] == before["fused"] + 1 and after["fallback"] == before["fallback"]
    assert after["materialized"] == before["materialized"], "a deferred tensor was computed eagerly"
    assert out["render"].shape == (3, H, W) and out["radii"].shape == (1, N)

    rcam = _ref_cam(cam, vc)
    m, v, o, b0 = ref.condition(scene, rcam)
    R = ref.rasterization_fwd(m, v, o, b0, scene.rgb, rcam.viewmat[None], rcam.K[None], W, H, backgrounds=bg[None])
    assert R["isect_ids"].numel() > 1000
    img = out["render"].permute(1, 2, 0)[None]
    torch.testing.assert_close(img, R["render_colors"], rtol=0, atol=IMG_ATOL)
    # integer work: bit-exact
    n_bad = int((out["radii"] != R["radii"]).sum())
    assert n_bad == 0, "radii differ for %d of %d primitives" % (n_bad, N)
    vis = R["radii"] > 0
    assert torch.equal(out["viewspace_points"][vis], R["means2d"][vis])

    g = torch.Generator(device="cuda").manual_seed(5)
    v_img = torch.randn(3, H, W, device="cuda", generator=g) / (H * W)
    (out["render"] * v_img).sum().backward()
    v_rc = v_img.permute(1, 2, 0)[None].contiguous()
    ref_grads, _ = ref.chain_grads(scene, rcam, bg, v_rc, torch.zeros(1, H, W, 1, device="cuda"))
    for nm, leaf, want in zip(("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"), model.leaves(), ref_grads):
        assert leaf.grad is not None, nm
        _assert_grad_close(nm, leaf.grad, want.reshape(leaf.grad.shape), rtol=3e-3)


def test_tile_lists_of_the_fused_route_are_bit_exact():
    ref = _ref()
    from ubs_b200 import synth

    D, N, W, H = 6, 120000, 800, 608
    scene = synth.make_scene(N, D, seed=11).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=12, device="cuda")[0]
    bg = torch.zeros(3, device="cuda")
    rc_mod, model = _caller(scene, bg, grad=False)
    vc = rc_mod.ViewpointCamera(cam)
    with torch.no_grad():
        means, convs, opacities = model.get_cond_mean_convariance_opacity(
            torch.nn.functional.normalize(model._xyz - cam.cam_pos[None], dim=-1))
        rgbs, alphas, meta = rc_mod.rasterization(
            means, None, None, opacities.squeeze(), model.get_beta[:, :1].squeeze(), model._rgb, cam.viewmat[None],
            vc.K()[None], W, H, backgrounds=bg[None], covars=convs)
    rcam = _ref_cam(cam, vc)
    q = torch.nn.functional.normalize(scene.xyz - cam.cam_pos[None], dim=-1)
    C_ = ref.load()
    scale, opacity, beta, mean = ref.activations(scene)
    ri, rj = ref.tril_rest(D, "cuda")
    rot = C_.l_triangle_to_rotmat_fwd(scene.l_triangle[:, :3].contiguous())
    covar = C_.rot_scale_l_triangle_to_covar_fwd(rot, scale.contiguous(), scene.l_triangle.contiguous(), ri, rj, False)
    m, v, o = C_.cond_mean_convariance_opacity_fwd(mean.contiguous(), covar, opacity.contiguous(),
                                                  beta[:, 1:].contiguous(), q.contiguous())
    R = ref.rasterization_fwd(m, v, o.squeeze(-1), beta[:, 0].contiguous(), scene.rgb, rcam.viewmat[None],
                              rcam.K[None], W, H, backgrounds=bg[None])
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert meta[k].shape == R[k].shape, (k, meta[k].shape, R[k].shape)
        assert torch.equal(meta[k], R[k]), "%s differs at %d entries" % (k, int((meta[k] != R[k]).sum()))
    vis = R["radii"] > 0
    assert torch.equal(meta["depths"][vis], R["depths"][vis])
    torch.testing.assert_close(meta["opacities"][vis], R["opacities"][vis], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(rgbs, R["render_colors"], rtol=0, atol=IMG_ATOL)
    torch.testing.assert_close(alphas, R["render_alphas"], rtol=0, atol=IMG_ATOL)


def test_batch_loop_renders_several_views_before_one_backward():
    """train.py:111-128: `batch_size` render() calls, the losses summed, ONE backward()."""
    ref = _ref()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import dropin, synth

    D, N, W, H, B = 6, 30000, 320, 240, 3
    scene = synth.make_scene(N, D, seed=21).to("cuda")
    cams = synth.make_cameras(B, W, H, seed=22, device="cuda")
    bg = torch.tensor([1.0, 1.0, 1.0], device="cuda")
    rc_mod, model = _caller(scene, bg, grad=True)
    g = torch.Generator(device="cuda").manual_seed(9)
    vs = [torch.randn(3, H, W, device="cuda", generator=g) / (H * W) for _ in range(B)]
    before = dropin.stats()
    total = 0.0
    images = []
    for cam, v in zip(cams, vs):
        out = model.render(rc_mod.ViewpointCamera(cam))
        images.append(out["render"])
        total = total + (out["render"] * v).sum()
    (total / B).backward()
    assert dropin.stats()["fused"] == before["fused"] + B
    want = None
    for cam, v, img in zip(cams, vs, images):
        rcam = _ref_cam(cam, rc_mod.ViewpointCamera(cam))
        grads, R = ref.chain_grads(scene, rcam, bg, (v / B).permute(1, 2, 0)[None].contiguous(),
                                   torch.zeros(1, H, W, 1, device="cuda"))
        torch.testing.assert_close(img.permute(1, 2, 0)[None], R["render_colors"], rtol=0, atol=IMG_ATOL)
        want = grads if want is None else [a + b.reshape(a.shape) for a, b in zip(want, grads)]
    for nm, leaf, w in zip(("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"), model.leaves(), want):
        _assert_grad_close(nm, leaf.grad, w.reshape(leaf.grad.shape), rtol=3e-3)


@pytest.mark.parametrize("mode", ["RGB", "Alpha", "Depth", "RGB+ED", "Normal"])
def test_viewer_call_with_quantile_mask_takes_the_fused_route(mode):
    """BetaModel.view (scene/beta_model.py:724-831): beta-quantile mask, GUI clip planes, every GUI render mode."""
    ref = _ref()
    from ubs_b200 import dropin, model as M, rendering, synth

    D, N, W, H = 7, 50000, 448, 336
    scene = synth.make_scene(N, D, seed=31).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=32, timestamps=[0.3], device="cuda")[0]
    bg = torch.tensor([0.0, 0.5, 1.0], device="cuda")
    rc_mod, model = _caller(scene, bg, grad=False)
    mask = M.quantile_mask(scene.beta, (10, 90), (0, 95), (5, 100))
    assert 0.4 * N < int(mask.sum()) < N
    c2w = torch.linalg.inv(cam.viewmat)
    vd = torch.nn.functional.normalize(scene.xyz - c2w[:3, 3][None], dim=-1)
    query = torch.cat([vd, torch.full((N, 1), 0.3, device="cuda")], dim=-1)
    before = dropin.stats()
    img, n_rendered = model.view_call(c2w, cam.K, W, H, query, mask, mode, 4.0, 11.0, 1.0)
    assert dropin.stats()["fused"] == before["fused"] + 1

    sub = synth.Scene(D, *[t[mask] for t in scene.tensors()])
    rcam = synth.Camera(torch.linalg.inv(c2w), cam.K, c2w[:3, 3].contiguous(), W, H, 0.3)
    mm, vv, oo, b0 = ref.condition(sub, rcam)
    kw = dict(near_plane=4.0, far_plane=11.0, radius_clip=1.0)
    R = ref.rasterization_fwd(mm, vv, oo, b0, sub.rgb, rcam.viewmat[None], cam.K[None], W, H, backgrounds=bg[None], **kw)
    assert n_rendered == int((R["radii"] > 0).sum()) and 0 < n_rendered < int(mask.sum())
    if mode == "RGB":
        torch.testing.assert_close(img, R["render_colors"], rtol=0, atol=IMG_ATOL)
        return
    if mode == "Alpha":
        torch.testing.assert_close(img, R["render_alphas"], rtol=0, atol=IMG_ATOL)
        return
    # the reference renders depth by compositing depths[..., None] as the colour over a zero background
    # (rendering.py:131-142); drive its kernel the same way (3 equal channels: the channel count it is built for)
    dcol = R["depths"][0][:, None].repeat(1, 3)
    Rd = ref.rasterization_fwd(mm, vv, oo, b0, dcol, rcam.viewmat[None], cam.K[None], W, H,
                               backgrounds=torch.zeros(1, 3, device="cuda"), **kw)
    depth_img = Rd["render_colors"][..., :1]
    tol = IMG_ATOL * max(float(depth_img.abs().max()), 1.0)
    if mode == "Depth":
        assert img.shape == (1, H, W, 1)
        torch.testing.assert_close(img, depth_img, rtol=0, atol=tol)
    elif mode == "RGB+ED":
        assert img.shape == (1, H, W, 4)
        torch.testing.assert_close(img[..., :3], R["render_colors"], rtol=0, atol=IMG_ATOL)
        ok = R["render_alphas"][..., 0] > 0.05
        want = depth_img / R["render_alphas"].clamp(min=1e-10)
        torch.testing.assert_close(img[..., 3:][ok], want[ok], rtol=1e-3, atol=1e-3)
    else:  # Normal: the reference's depth_to_normal on the reference depth image, mapped to [0, 1]
        from test_gpu_pinned import _well_conditioned_normals

        want = (rendering.depth_to_normal(depth_img, c2w[None], cam.K[None]) + 1) / 2
        assert img.shape == (1, H, W, 3)
        # compare where the normal is not the direction of a near-zero vector
        well = _well_conditioned_normals(depth_img, c2w[None], cam.K[None])
        assert well.float().mean() > 0.05
        assert ((img - want).abs().max(dim=-1).values[well] < 2e-2).float().mean() > 0.98


def test_flatten_ids_of_a_filtered_view_index_the_kept_primitives():
    ref = _ref()
    from ubs_b200 import synth

    D, N, W, H = 6, 20000, 256, 192
    scene = synth.make_scene(N, D, seed=41).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=42, device="cuda")[0]
    bg = torch.zeros(3, device="cuda")
    rc_mod, model = _caller(scene, bg, grad=False)
    mask = torch.rand(N, device="cuda") < 0.6
    q = torch.nn.functional.normalize(scene.xyz - cam.cam_pos[None], dim=-1)
    with torch.no_grad():
        means, convs, opacities = model.get_cond_mean_convariance_opacity(q)
        _, _, meta = rc_mod.rasterization(means[mask], None, None, opacities.squeeze()[mask],
                                          model.get_beta[:, :1].squeeze()[mask], model._rgb[mask], cam.viewmat[None],
                                          cam.K[None], W, H, backgrounds=bg[None], covars=convs[mask])
    sub = synth.Scene(D, *[t[mask] for t in scene.tensors()])
    mm, vv, oo, b0 = ref.condition(sub, cam)
    R = ref.rasterization_fwd(mm, vv, oo, b0, sub.rgb, cam.viewmat[None], cam.K[None], W, H, backgrounds=bg[None])
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert meta[k].shape == R[k].shape and torch.equal(meta[k], R[k]), k


def test_any_other_use_of_the_deferred_tensors_falls_back_to_the_operator_chain():
    ref = _ref()
    from ubs_b200 import dropin, synth

    D, N, W, H = 6, 20000, 256, 192
    scene = synth.make_scene(N, D, seed=51).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=52, device="cuda")[0]
    bg = torch.tensor([0.3, 0.3, 0.3], device="cuda")
    rc_mod, model = _caller(scene, bg, grad=True)
    q = torch.nn.functional.normalize(scene.xyz - cam.cam_pos[None], dim=-1)
    means, convs, opacities = model.get_cond_mean_convariance_opacity(q)
    before = dropin.stats()
    shifted = means + 0.0  # arithmetic on a deferred tensor: it is computed by the stand-alone operators
    assert not isinstance(shifted, dropin.Deferred) and shifted.requires_grad
    rgbs, alphas, meta = rc_mod.rasterization(shifted, None, None, opacities.squeeze(), model.get_beta[:, 0], model._rgb,
                                              cam.viewmat[None], cam.K[None], W, H, backgrounds=bg[None], covars=convs)
    after = dropin.stats()
    assert after["fallback"] == before["fallback"] + 1 and after["fused"] == before["fused"]
    m, v, o, b0 = ref.condition(scene, cam)
    R = ref.rasterization_fwd(m, v, o, b0, scene.rgb, cam.viewmat[None], cam.K[None], W, H, backgrounds=bg[None])
    torch.testing.assert_close(rgbs, R["render_colors"], rtol=0, atol=IMG_ATOL)
    rgbs.sum().backward()
    assert model._scale.grad is not None and float(model._scale.grad.abs().sum()) > 0
    # get_xyz_covariance (train.py:154, the SGLD noise) is the spatial block: eager, same values as the operator
    from ubs_b200 import ops

    with torch.no_grad():
        xyz_cov = model.get_xyz_covariance
        want = ops.rot_scale_l_triangle_to_covar(ops.l_triangle_to_rotmat(scene.l_triangle[:, :3].contiguous()),
                                                 torch.nn.functional.softplus(scene.scale), scene.l_triangle,
                                                 model.rest_i, model.rest_j, spatial_block=True)
    assert not isinstance(xyz_cov, dropin.Deferred) and torch.equal(xyz_cov, want)
