"""GPU parity of the forward stages against the reference's own CUDA kernels (oracle/_ref) -- stage-isolated.

Bar: integer outputs (radii, tile counts, keys, sorted order, offsets) bit-exact; floats within the stated
tolerance (north star: max-abs 1e-4 on images).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

IMG_ATOL = 1e-4  # north-star tolerance on rendered images


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    ref_cuda.load()
    return ref_cuda


def _conditioned_inputs(N, seed, W, H, n_cams=1, device="cuda"):
    from ubs_b200 import synth

    g = torch.Generator().manual_seed(seed)
    means = ((torch.rand(N, 3, generator=g) * 2 - 1) * 4.0).to(device)
    covars = synth.random_spd_covars(N, seed + 1, 0.01, 0.3, device=device)
    opac = torch.rand(N, generator=g).to(device)
    betas = (4.0 * torch.exp(torch.randn(N, generator=g) * 0.5)).to(device)
    colors = torch.rand(N, 3, generator=g).to(device)
    cams = synth.make_cameras(n_cams, W, H, radius=8.0, seed=seed)
    viewmats = torch.stack([c.viewmat for c in cams]).to(device)
    Ks = torch.stack([c.K for c in cams]).to(device)
    return means, covars, opac, betas, colors, viewmats, Ks


@pytest.mark.parametrize("N,W,H,C", [(20000, 320, 240, 1), (50000, 800, 800, 1), (8000, 200, 136, 3)])
def test_projection_fwd_matches_reference(N, W, H, C):
    ref = _ref()
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 11 + N, W, H, C)
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    cov6 = covars[..., tri[0], tri[1]].contiguous()
    for comp in (False, True):
        r_radii, r_m2d, r_depth, r_conic, r_comp = ref.load().fully_fused_projection_fwd(
            means, cov6, None, None, viewmats, Ks, W, H, 0.3, 0.01, 1e10, 0.0, comp, False)
        radii, m2d, depth, conic, comps = ops.projection_fwd(means, cov6, viewmats, Ks, W, H, 0.3, 0.01, 1e10, 0.0, comp)
        assert torch.equal(radii, r_radii), "radii differ at %d entries" % int((radii != r_radii).sum())
        vis = r_radii > 0
        assert vis.sum() > 0
        assert torch.equal(depth[vis], r_depth[vis]), "depth bits differ"
        assert torch.equal(m2d[vis], r_m2d[vis]), "means2d bits differ"
        torch.testing.assert_close(conic[vis], r_conic[vis], rtol=1e-4, atol=1e-6)
        if comp:
            torch.testing.assert_close(comps[vis], r_comp[vis], rtol=1e-5, atol=1e-6)
        # culled entries are zeroed by us (uninitialised in the reference)
        assert (m2d[~vis] == 0).all() and (depth[~vis] == 0).all()


def test_projection_culling_knobs():
    ref = _ref()
    from ubs_b200 import ops

    N, W, H = 30000, 400, 300
    means, covars, *_rest, viewmats, Ks = _conditioned_inputs(N, 5, W, H, 1)
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    cov6 = covars[..., tri[0], tri[1]].contiguous()
    for near, far, clip in [(6.0, 9.0, 0.0), (0.01, 1e10, 3.0), (7.5, 8.5, 2.0)]:
        r = ref.load().fully_fused_projection_fwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3, near, far, clip,
                                                  False, False)
        o = ops.projection_fwd(means, cov6, viewmats, Ks, W, H, 0.3, near, far, clip, False)
        assert torch.equal(o[0], r[0])


@pytest.mark.parametrize("n,bits", [(0, 40), (1, 64), (1000, 46), (4096, 33), (4097, 46), (300001, 52), (2_000_003, 46)])
def test_radix_sort_is_stable_and_sorted(n, bits):
    from ubs_b200 import ops

    g = torch.Generator().manual_seed(n + bits)
    # few distinct keys in the upper bits => many ties, exercising stability
    hi = torch.randint(0, 97, (n,), generator=g, dtype=torch.int64) << 32
    lo = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int64) << 11
    keys = ((hi | lo) & ((1 << bits) - 1 if bits < 64 else -1)).cuda()
    vals = torch.arange(n, dtype=torch.int32).cuda()
    k2, v2 = ops.radix_sort_pairs(keys, vals, 0, bits)
    ks, order = torch.sort(keys, stable=True)
    assert torch.equal(k2, ks)
    assert torch.equal(v2.long(), order)


def test_radix_sort_partial_bit_range():
    from ubs_b200 import ops

    g = torch.Generator().manual_seed(3)
    n = 50000
    keys = torch.randint(0, 1 << 62, (n,), generator=g, dtype=torch.int64).cuda()
    vals = torch.arange(n, dtype=torch.int32).cuda()
    k2, v2 = ops.radix_sort_pairs(keys, vals, 8, 29)
    sub = (keys >> 8) & ((1 << 21) - 1)
    _, order = torch.sort(sub, stable=True)
    assert torch.equal(v2.long(), order)
    assert torch.equal(k2, keys[order])


@pytest.mark.parametrize("N,W,H,C", [(20000, 320, 240, 1), (60000, 800, 800, 1), (8000, 200, 136, 3),
                                      (3000, 1920, 1080, 1)])
def test_isect_sort_offsets_bit_exact(N, W, H, C):
    ref = _ref()
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 101 + N, W, H, C)
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, flat, offs = ops.isect_tiles(R["means2d"], R["radii"], R["depths"], 16, tw, th, n_cameras=C,
                                           return_offsets=True)
    assert torch.equal(tpg, R["tiles_per_gauss"])
    assert ids.shape == R["isect_ids"].shape
    assert torch.equal(ids, R["isect_ids"])
    assert torch.equal(flat, R["flatten_ids"])
    assert torch.equal(offs, R["isect_offsets"])
    assert torch.equal(ops.isect_offset_encode(R["isect_ids"], C, tw, th), R["isect_offsets"])
    # unsorted emission order is also the reference's
    tpg_u, ids_u, flat_u = ops.isect_tiles(R["means2d"], R["radii"], R["depths"], 16, tw, th, sort=False, n_cameras=C)
    r_u = ref.load().isect_tiles(R["means2d"], R["radii"], R["depths"], None, None, C, 16, tw, th, False, True)
    assert torch.equal(ids_u, r_u[1]) and torch.equal(flat_u, r_u[2])


def test_isect_empty_and_all_culled():
    from ubs_b200 import ops

    C, N, tw, th = 2, 100, 5, 4
    m2d = torch.zeros(C, N, 2, device="cuda")
    radii = torch.zeros(C, N, dtype=torch.int32, device="cuda")
    depths = torch.ones(C, N, device="cuda")
    tpg, ids, flat, offs = ops.isect_tiles(m2d, radii, depths, 16, tw, th, n_cameras=C, return_offsets=True)
    assert ids.numel() == 0 and flat.numel() == 0
    assert (tpg == 0).all() and (offs == 0).all()
    assert (ops.isect_offset_encode(ids, C, tw, th) == 0).all()


@pytest.mark.parametrize("N,W,H,C,bg", [(20000, 320, 240, 1, True), (60000, 800, 800, 1, False),
                                         (8000, 200, 136, 3, True), (40000, 333, 207, 1, True)])
def test_rasterize_fwd_matches_reference(N, W, H, C, bg):
    ref = _ref()
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 977 + N, W, H, C)
    opac = opac * 0.9 + 0.1
    backgrounds = torch.rand(C, 3, device="cuda") if bg else None
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=backgrounds)
    rc, ra, last = ops.rasterize_fwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], backgrounds,
                                     None, W, H, 16, R["isect_offsets"], R["flatten_ids"])
    assert (R["render_alphas"] > 0.5).float().mean() > 0.05, "scene too empty to be a meaningful test"
    torch.testing.assert_close(rc, R["render_colors"], rtol=0, atol=IMG_ATOL)
    torch.testing.assert_close(ra, R["render_alphas"], rtol=0, atol=IMG_ATOL)
    match = (last == R["last_ids"]).float().mean().item()
    assert match > 0.999, "last_ids match rate %.5f" % match


@pytest.mark.parametrize("mode", ["RGB", "RGB+D", "RGB+ED", "Depth", "EDepth"])
def test_rasterization_end_to_end(mode):
    ref = _ref()
    import ubs_b200

    N, W, H, C = 50000, 640, 480, 1
    means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 4242, W, H, C)
    bg = torch.tensor([[0.2, 0.4, 0.6]], device="cuda")
    rc, ra, meta = ubs_b200.rasterization(means, None, None, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg,
                                          render_mode=mode, covars=covars)
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(meta[k], R[k]), k
    assert torch.equal(meta["depths"][R["radii"] > 0], R["depths"][R["radii"] > 0])
    torch.testing.assert_close(ra, R["render_alphas"], rtol=0, atol=IMG_ATOL)
    if mode == "RGB":
        torch.testing.assert_close(rc, R["render_colors"], rtol=0, atol=IMG_ATOL)
        assert rc.shape == (C, H, W, 3)
    elif mode in ("RGB+D", "RGB+ED"):
        assert rc.shape == (C, H, W, 4)
        torch.testing.assert_close(rc[..., :3], R["render_colors"], rtol=0, atol=IMG_ATOL)
    else:
        assert rc.shape == (C, H, W, 1)
    assert set(meta) >= {"camera_ids", "primitive_ids", "radii", "means2d", "depths", "conics", "opacities", "betas",
                         "tile_width", "tile_height", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets",
                         "width", "height", "tile_size", "n_cameras"}
