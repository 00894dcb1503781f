"""GPU parity of the CUDA library against (1) the committed fixtures produced by the reference's CUDA kernels
(tests/golden/ref_cuda_D{6,7}.npz) -- needs neither /root/reference nor oracle/_ref -- and (2) the CPU oracle on the
same seeded inputs.  Integers bit-exact, images 1e-4 abs, gradients 1e-3 of scale per stage (3e-3 end to end)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _load(D):
    z = np.load(os.path.join(HERE, "golden", "ref_cuda_D%d.npz" % D))
    return {k: torch.from_numpy(z[k]).cuda() for k in z.files}


def _grad_close(name, mine, theirs, rtol=1e-3):
    scale = theirs.abs().max().clamp_min(1e-20)
    err = ((mine - theirs).abs().max() / scale).item()
    assert err < rtol, "%s: max abs err / max |ref| = %.3e" % (name, err)


@pytest.mark.parametrize("D", [6, 7])
def test_stages_against_reference_fixture(D):
    from make_golden_ref_cuda import SCENES
    from ubs_b200 import ops

    g = _load(D)
    W, H = SCENES[D]["W"], SCENES[D]["H"]
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, flat, off = ops.isect_tiles(g["fwd_means2d"], g["fwd_radii"], g["fwd_depths"], 16, tw, th, n_cameras=1,
                                          return_offsets=True)
    assert torch.equal(tpg, g["fwd_tiles_per_gauss"])
    assert torch.equal(ids, g["fwd_isect_ids"])
    assert torch.equal(flat, g["fwd_flatten_ids"])
    assert torch.equal(off, g["fwd_isect_offsets"])
    bg = torch.tensor([[0.2, 0.5, 0.9]], device="cuda")
    rc, ra, last = ops.rasterize_fwd(g["fwd_means2d"], g["fwd_conics"], g["fwd_colors"], g["fwd_opacities"],
                                     g["fwd_betas"], bg, None, W, H, 16, off, flat)
    torch.testing.assert_close(rc, g["fwd_render_colors"], rtol=0, atol=1e-4)
    torch.testing.assert_close(ra, g["fwd_render_alphas"], rtol=0, atol=1e-4)
    assert (last == g["fwd_last_ids"]).float().mean() > 0.9995
    from make_golden_ref_cuda import scene_and_camera

    _, _, _, v_rc, v_ra = scene_and_camera(D, "cuda")
    grads = ops.rasterize_bwd(g["fwd_means2d"], g["fwd_conics"], g["fwd_colors"], g["fwd_opacities"], g["fwd_betas"],
                              bg, None, W, H, 16, off, flat, g["fwd_render_alphas"], g["fwd_last_ids"], v_rc, v_ra)
    for name, a in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), grads):
        _grad_close(name, a, g["mid_" + name])


@pytest.mark.parametrize("D", [6, 7])
def test_fused_path_against_reference_fixture(D):
    from make_golden_ref_cuda import scene_and_camera
    from ubs_b200 import fused

    g = _load(D)
    scene, cam, bg, v_rc, v_ra = scene_and_camera(D, "cuda")
    rec = fused.pack_records(D, *scene.tensors()).requires_grad_(True)
    rz = fused.FusedRasterizer(D, scene.N, cam.width, cam.height, n_cams=1)
    ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
    rc, ra = fused.render(rec, rz, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    assert (rz.radii == g["fwd_radii"]).float().mean() > 0.999
    assert abs(rz.last_pair_count() - g["fwd_isect_ids"].numel()) <= 8
    torch.testing.assert_close(rc, g["fwd_render_colors"], rtol=0, atol=1e-4)
    torch.testing.assert_close(ra, g["fwd_render_alphas"], rtol=0, atol=1e-4)
    torch.autograd.backward((rc, ra), (v_rc, v_ra))
    for name, mine in zip(("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"),
                          fused.unpack_records(D, rec.grad)):
        _grad_close(name, mine, g["grad_" + name].reshape(mine.shape), rtol=3e-3)


@pytest.mark.parametrize("N,W,H,C", [(5000, 200, 120, 2), (0, 64, 48, 1), (1, 33, 17, 1)])
def test_cuda_tile_lists_and_compositing_equal_cpu_oracle(N, W, H, C):
    """Same seeded inputs through the C-ABI and through oracle/raster_oracle.c, including the empty and
    single-primitive edge cases and image sizes that are not multiples of the tile."""
    from oracle import ubs_oracle as O
    from ubs_b200 import ops

    g = torch.Generator().manual_seed(N + W)
    m2d = torch.rand(C, N, 2, generator=g) * torch.tensor([W + 40.0, H + 40.0]) - 20.0
    radii = torch.randint(0, 30, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g) * 10 + 0.1
    depths[:, : N // 10] = depths[:, :1]  # equal depths: stability of the sort decides the order
    s = 0.002 + 0.05 * torch.rand(C, N, generator=g)
    conics = torch.stack([s, (torch.rand(C, N, generator=g) - 0.5) * 0.5 * s, s * (0.5 + torch.rand(C, N, generator=g))], -1)
    colors = torch.rand(C, N, 3, generator=g)
    opac = torch.rand(C, N, generator=g)
    betas = 0.5 + 4 * torch.rand(C, N, generator=g)
    bg = torch.rand(C, 3, generator=g)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg_o, ids_o, flat_o = O.isect_tiles(m2d, radii, depths, 16, tw, th)
    off_o = O.isect_offset_encode(ids_o, C, tw, th)
    d = lambda t: t.cuda()  # noqa: E731
    tpg, ids, flat, off = ops.isect_tiles(d(m2d), d(radii), d(depths), 16, tw, th, n_cameras=C, return_offsets=True)
    assert torch.equal(tpg.cpu(), tpg_o) and torch.equal(ids.cpu(), ids_o) and torch.equal(flat.cpu(), flat_o)
    assert torch.equal(off.cpu(), off_o)
    rc_o, ra_o, li_o = O.rasterize_fwd(m2d, conics, colors, opac, betas, bg, None, W, H, 16, off_o, flat_o)
    rc, ra, li = ops.rasterize_fwd(d(m2d), d(conics), d(colors), d(opac), d(betas), d(bg), None, W, H, 16, off, flat)
    assert (rc.cpu() - rc_o).abs().max() < 2e-4 and (ra.cpu() - ra_o).abs().max() < 2e-4
    if N > 0:
        assert (li.cpu() == li_o).float().mean() > 0.999
    v_rc = torch.randn(C, H, W, 3, generator=g) / (H * W)
    v_ra = torch.randn(C, H, W, 1, generator=g) / (H * W)
    g_o = O.rasterize_bwd(m2d, conics, colors, opac, betas, bg, None, W, H, 16, off_o, flat_o, ra.cpu(), li.cpu(), v_rc,
                          v_ra)
    g_c = ops.rasterize_bwd(d(m2d), d(conics), d(colors), d(opac), d(betas), d(bg), None, W, H, 16, off, flat, ra, li,
                            d(v_rc), d(v_ra))
    for name, a, b in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_betas"), g_c, g_o):
        if N > 0 and b.abs().max() > 0:
            _grad_close(name, a.cpu(), b)
        else:
            assert a.abs().max().item() == 0 if a.numel() else True
