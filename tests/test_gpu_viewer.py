"""GPU parity of the viewer-path options: beta-quantile primitive mask, near/far planes,
radius_clip, Alpha and Depth modes of BetaModel.view (scene/beta_model.py:724-831) against the reference CUDA
kernels driven on the masked subset of primitives, as the reference does (`means[mask]`, ...)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
IMG_ATOL = 1e-4


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _subset(scene, mask):
    from ubs_b200 import synth

    return synth.Scene(scene.D, *[t[mask] for t in scene.tensors()])


@pytest.mark.parametrize("D", [6, 7])
def test_view_rgb_with_quantile_mask_and_clip_planes(D):
    ref = _ref()
    from ubs_b200 import fused, model, synth

    N, W, H = 50000, 512, 384
    scene = synth.make_scene(N, D, seed=99 + D).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=4, timestamps=[0.6], device="cuda")[0]
    m = model.PackedBetaModel(fused.pack_records(D, *scene.tensors()), D)
    c2w = torch.linalg.inv(cam.viewmat)
    opts = dict(b_xyz=(10, 85), b_view=(5, 100), b_time=(0, 90), timestamp=0.6, near_plane=5.0, far_plane=9.5,
                radius_clip=1.5, backgrounds=(255, 128, 0))
    img, n_rendered = m.view(c2w, cam.K, W, H, "RGB", **opts)
    mask = model.quantile_mask(scene.beta, (10, 85), (5, 100), (0, 90) if D == 7 else None)
    assert 0.3 * N < int(mask.sum()) < N
    sub = _subset(scene, mask)
    cam_ref = synth.Camera(torch.linalg.inv(c2w), cam.K, c2w[:3, 3].contiguous(), W, H, 0.6)
    mm, vv, oo, b0 = ref.condition(sub, cam_ref)
    bg = torch.tensor([1.0, 128 / 255.0, 0.0], device="cuda")
    R = ref.rasterization_fwd(mm, vv, oo, b0, sub.rgb, cam_ref.viewmat[None], cam.K[None], W, H, backgrounds=bg[None],
                              near_plane=5.0, far_plane=9.5, radius_clip=1.5)
    assert n_rendered == int((R["radii"] > 0).sum())
    assert 0 < n_rendered < int(mask.sum())  # the planes and the clip really cull
    torch.testing.assert_close(img, R["render_colors"][0], rtol=0, atol=IMG_ATOL)
    alpha, _ = m.view(c2w, cam.K, W, H, "Alpha", **opts)
    torch.testing.assert_close(alpha, R["render_alphas"][0], rtol=0, atol=IMG_ATOL)


def test_view_depth_modes_match_reference_chain():
    ref = _ref()
    from ubs_b200 import fused, model, synth

    D, N, W, H = 6, 40000, 400, 300
    scene = synth.make_scene(N, D, seed=5).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=8, device="cuda")[0]
    m = model.PackedBetaModel(fused.pack_records(D, *scene.tensors()), D)
    c2w = torch.linalg.inv(cam.viewmat)
    mask = model.quantile_mask(scene.beta, (0, 70), (0, 100), None)
    sub = _subset(scene, mask)
    mm, vv, oo, b0 = ref.condition(sub, cam)
    # the reference renders depth by passing depths[..., None] as the colour (rendering.py:131-142)
    P = ref.rasterization_fwd(mm, vv, oo, b0, sub.rgb, cam.viewmat[None], cam.K[None], W, H)
    Rd = ref.rasterization_fwd(mm, vv, oo, b0, P["depths"][0][:, None].repeat(1, 3), cam.viewmat[None], cam.K[None],
                               W, H, backgrounds=torch.zeros(1, 3, device="cuda"))
    depth, n = m.view(c2w, cam.K, W, H, "Depth", b_xyz=(0, 70))
    assert depth.shape == (H, W, 1) and n == int((P["radii"] > 0).sum())
    scale = float(Rd["render_colors"].abs().max())
    torch.testing.assert_close(depth, Rd["render_colors"][0, ..., :1], rtol=0, atol=IMG_ATOL * max(scale, 1.0))
    ed, _ = m.view(c2w, cam.K, W, H, "RGB+ED", b_xyz=(0, 70))
    assert ed.shape == (H, W, 4)
    want = Rd["render_colors"][0, ..., :1] / Rd["render_alphas"][0].clamp(min=1e-10)
    ok = Rd["render_alphas"][0, ..., 0] > 0.05  # the normalisation amplifies rounding where alpha is tiny
    torch.testing.assert_close(ed[..., 3:][ok], want[ok], rtol=1e-3, atol=1e-3)


def test_render_returns_reference_dict_and_supports_mask():
    from ubs_b200 import fused, model, synth

    D, N, W, H = 6, 20000, 256, 192
    scene = synth.make_scene(N, D, seed=6).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=1, device="cuda")[0]
    m = model.PackedBetaModel(fused.pack_records(D, *scene.tensors()), D, background=torch.ones(3, device="cuda"))
    out = m.render(cam)
    assert set(out) >= {"render", "viewspace_points", "visibility_filter", "radii", "is_used"}
    assert out["render"].shape == (3, H, W)
    keep = torch.rand(N, device="cuda") < 0.5
    out2 = m.render(cam, mask=keep)
    assert not bool((out2["radii"][0][~keep] > 0).any())
    assert torch.equal(out2["radii"][0][keep], out["radii"][0][keep])
