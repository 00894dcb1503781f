"""GPU parity of the tile-binning + per-tile segment sort route (csrc/bin_sort.cu) against (a) the reference's own
isect_tiles / isect_offset_encode CUDA kernels and (b) a torch stable sort of the reference's key definition
(isect_tiles.cu:82-95), which needs no reference build.  Bar: bit-exact (integer work)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _expected(means2d, radii, depths, tile_size, tw, th):
    """Sorted pair list by the reference's definition: key = cam << (32+tb) | tile << 32 | depth bits, val =
    cam*N + prim, emitted in (cam, prim, tile row-major) order, stable ascending sort (isect_tiles.cu:39-96,230-278)."""
    C, N, _ = means2d.shape
    ts = float(tile_size)
    r = radii.float() / ts
    t = means2d / ts
    x0 = torch.floor(t[..., 0] - r).clamp(0, tw).int()
    y0 = torch.floor(t[..., 1] - r).clamp(0, th).int()
    x1 = torch.ceil(t[..., 0] + r).clamp(0, tw).int()
    y1 = torch.ceil(t[..., 1] + r).clamp(0, th).int()
    vis = radii > 0
    tpg = torch.where(vis, (x1 - x0) * (y1 - y0), torch.zeros_like(x0))
    n_tiles = tw * th
    tb = max(1, int(n_tiles).bit_length())
    keys, vals = [], []
    idx = torch.nonzero(vis.flatten()).flatten()
    x0f, y0f, x1f, y1f = (a.flatten()[idx].long() for a in (x0, y0, x1, y1))
    dbits = depths.flatten()[idx].view(torch.int32).long()
    w, h = x1f - x0f, y1f - y0f
    cnt = w * h
    rep = torch.repeat_interleave(torch.arange(idx.numel(), device=idx.device), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(rep.numel(), device=idx.device) - start[rep]
    ty = y0f[rep] + local // w[rep].clamp_min(1)
    tx = x0f[rep] + local % w[rep].clamp_min(1)
    cam = idx[rep] // N
    key = (cam << (32 + tb)) | ((ty * tw + tx) << 32) | dbits[rep]
    val = idx[rep].int()
    key_s, order = torch.sort(key, stable=True)
    val_s = val[order]
    slot = (key_s >> 32 >> tb) * n_tiles + ((key_s >> 32) & ((1 << tb) - 1))
    offs = torch.searchsorted(slot, torch.arange(C * n_tiles, device=idx.device)).int().view(C, th, tw)
    if key_s.numel() == 0:
        offs.zero_()
    return tpg.int(), key_s, val_s, offs


def _random_splats(C, N, W, H, seed, rmax=40, tie_frac=0.0, depth_lo=0.5, depth_hi=60.0):
    g = torch.Generator().manual_seed(seed)
    means2d = torch.rand(C, N, 2, generator=g) * torch.tensor([W + 60.0, H + 60.0]) - 30.0
    radii = torch.randint(0, rmax, (C, N), generator=g, dtype=torch.int32)
    radii[torch.rand(C, N, generator=g) < 0.3] = 0
    depths = depth_lo + (depth_hi - depth_lo) * torch.rand(C, N, generator=g)
    if tie_frac > 0:
        # many bit-identical depths: the order inside a run must be the flatten id
        pool = depth_lo + (depth_hi - depth_lo) * torch.rand(7, generator=g)
        m = torch.rand(C, N, generator=g) < tie_frac
        depths[m] = pool[torch.randint(0, 7, (int(m.sum()),), generator=g)]
    return means2d.cuda(), radii.cuda(), depths.cuda()


@pytest.mark.parametrize("C,N,W,H,rmax,tie", [(1, 30000, 640, 480, 40, 0.0), (1, 30000, 640, 480, 40, 0.5),
                                               (3, 9000, 200, 136, 30, 0.2), (1, 500, 1920, 1080, 20, 0.0),
                                               (2, 4000, 64, 48, 200, 0.3),   # every tile far above kSegMax
                                               (1, 3000, 96, 96, 60, 1.0),    # all depths from a pool of 7
                                               # tiles of 2049..4096 / 4097..8192 pairs (the 16- and 32-pairs-per-thread
                                               # shared-memory sorts) mixed with shorter and longer ones
                                               (1, 40000, 320, 240, 50, 0.0), (1, 60000, 320, 240, 60, 0.3),
                                               (2, 25000, 400, 300, 70, 0.1)])
def test_bin_sort_matches_stable_sort_definition(C, N, W, H, rmax, tie):
    from ubs_b200 import ops

    means2d, radii, depths = _random_splats(C, N, W, H, 17 + N + C, rmax=rmax, tie_frac=tie)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    e_tpg, e_ids, e_flat, e_offs = _expected(means2d, radii, depths, 16, tw, th)
    tpg, ids, flat, offs = ops.isect_tiles(means2d, radii, depths, 16, tw, th, n_cameras=C, return_offsets=True,
                                           method="bin")
    assert torch.equal(tpg, e_tpg)
    assert ids.shape == e_ids.shape
    assert torch.equal(offs, e_offs)
    assert torch.equal(ids, e_ids)
    assert torch.equal(flat, e_flat)
    # and the onesweep route agrees with both
    tpg2, ids2, flat2, offs2 = ops.isect_tiles(means2d, radii, depths, 16, tw, th, n_cameras=C, return_offsets=True)
    assert torch.equal(ids2, ids) and torch.equal(flat2, flat) and torch.equal(offs2, offs) and torch.equal(tpg2, tpg)


def test_bin_sort_empty_and_all_culled():
    from ubs_b200 import ops

    C, N, tw, th = 2, 100, 5, 4
    m2d = torch.zeros(C, N, 2, device="cuda")
    radii = torch.zeros(C, N, dtype=torch.int32, device="cuda")
    depths = torch.ones(C, N, device="cuda")
    tpg, ids, flat, offs = ops.isect_tiles(m2d, radii, depths, 16, tw, th, n_cameras=C, return_offsets=True,
                                           method="bin")
    assert ids.numel() == 0 and flat.numel() == 0 and (tpg == 0).all() and (offs == 0).all()
    tpg, ids, flat, offs = ops.isect_tiles(m2d[:, :0], radii[:, :0], depths[:, :0], 16, tw, th, n_cameras=C,
                                           return_offsets=True, method="bin")
    assert ids.numel() == 0 and (offs == 0).all()


@pytest.mark.parametrize("N,W,H,C", [(20000, 320, 240, 1), (60000, 800, 800, 1), (8000, 200, 136, 3)])
def test_bin_sort_matches_reference_kernels(N, W, H, C):
    from oracle import ref_cuda as ref

    if not ref.available():
        pytest.skip("reference CUDA oracle not built")
    from test_gpu_forward_stages import _conditioned_inputs
    from ubs_b200 import ops

    means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 101 + N, W, H, C)
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, flat, offs = ops.isect_tiles(R["means2d"], R["radii"], R["depths"], 16, tw, th, n_cameras=C,
                                           return_offsets=True, method="bin")
    assert torch.equal(tpg, R["tiles_per_gauss"])
    assert torch.equal(ids, R["isect_ids"])
    assert torch.equal(flat, R["flatten_ids"])
    assert torch.equal(offs, R["isect_offsets"])


@pytest.mark.parametrize("D,N,W,H,C", [(6, 50000, 640, 400, 1), (7, 30000, 320, 256, 2)])
def test_fused_rasterizer_bin_equals_onesweep(D, N, W, H, C):
    """The two sort routes of the fused path give identical lists, offsets and images; an undersized capacity is
    reported, clamps the offsets and grows on the next call."""
    from ubs_b200 import fused, synth

    scene = synth.make_scene(N, D, seed=5 + D).to("cuda")
    cams = synth.make_cameras(C, W, H, seed=2, timestamps=[0.3, 0.8][:C])
    rec = fused.pack_records(D, *scene.tensors())
    V = torch.stack([c.viewmat for c in cams]).cuda()
    K = torch.stack([c.K for c in cams]).cuda()
    P = torch.stack([c.cam_pos for c in cams]).cuda()
    ts = torch.tensor([c.timestamp for c in cams], device="cuda") if D == 7 else None
    bg = torch.rand(C, 3, device="cuda")
    out = {}
    for mode in ("bin", "onesweep"):
        rz = fused.FusedRasterizer(D, N, W, H, n_cams=C, sort_mode=mode)
        rc, ra = rz.forward(rec, V, K, P, ts, bg)
        n = rz.last_pair_count()
        out[mode] = (n, rz.isect_ids[:n].clone(), rz.flatten_ids[:n].clone(), rz.offsets.clone(), rc.clone(),
                     ra.clone(), rz.last_ids.clone())
        assert not rz.overflowed()
    assert out["bin"][0] == out["onesweep"][0] > 0
    for a, b in zip(out["bin"][1:], out["onesweep"][1:]):
        assert torch.equal(a, b)
    n = out["bin"][0]
    small = fused.FusedRasterizer(D, N, W, H, n_cams=C, capacity=n // 3, sort_mode="bin")
    small.forward(rec, V, K, P, ts, bg)
    assert small.overflowed() and small.last_pair_count() == n
    assert int(small.offsets.max()) <= n // 3
    torch.cuda.synchronize()
    small.forward(rec, V, K, P, ts, bg)  # polls the count, grows
    rc, ra = small.forward(rec, V, K, P, ts, bg)
    assert small.capacity >= n and not small.overflowed()
    assert torch.equal(small.isect_ids[:n], out["bin"][1]) and torch.equal(rc, out["bin"][4])
