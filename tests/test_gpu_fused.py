"""GPU parity of the fused fast path (packed records -> image) against the reference kernel chain
K1 -> K2 -> K3 -> K5 -> K7-K9 -> K10 driven exactly as scene/beta_model.py:660-711 drives it."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
IMG_ATOL = 1e-4


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _reference_frame(ref, scene, cam, bg):
    m, v, o, b0 = ref.condition(scene, cam)
    return ref.rasterization_fwd(m, v, o, b0, scene.rgb, cam.viewmat[None], cam.K[None], cam.width, cam.height,
                                 backgrounds=bg[None])


@pytest.mark.parametrize("D,N,W,H", [(6, 60000, 640, 480), (7, 40000, 507, 380), (6, 200000, 800, 800)])
def test_fused_forward_matches_reference_chain(D, N, W, H):
    ref = _ref()
    from ubs_b200 import fused, synth

    scene = synth.make_scene(N, D, seed=1234 + D).to("cuda")
    cams = synth.make_cameras(2, W, H, seed=9, timestamps=[0.25, 0.8], device="cuda")
    bg = torch.tensor([0.1, 0.3, 0.5], device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    for cam in cams:
        R = _reference_frame(ref, scene, cam, bg)
        ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
        rc, ra = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
        assert not rz.overflowed()
        n_ref = R["isect_ids"].numel()
        assert n_ref > 1000
        vis = R["radii"] > 0
        # integer-deciding quantities are bit-exact: the fused kernel evaluates the activations and the view direction
        # with the same roundings as torch does for the reference chain
        radii_match = (rz.radii == R["radii"]).float().mean().item()
        depth_match = (rz.depths[vis] == R["depths"][vis]).float().mean().item()
        n_radii_bad, n_depth_bad = int((rz.radii != R["radii"]).sum()), int((rz.depths[vis] != R["depths"][vis]).sum())
        assert n_radii_bad == 0 and n_depth_bad == 0, (n_radii_bad, n_depth_bad)
        assert torch.equal(rz.means2d[vis], R["means2d"][vis])
        n = rz.last_pair_count()
        assert n == n_ref
        assert torch.equal(rz.isect_ids[:n], R["isect_ids"]) and torch.equal(rz.flatten_ids[:n], R["flatten_ids"])
        assert torch.equal(rz.offsets, R["isect_offsets"])
        print("D=%d radii exact %.5f depth bit-exact %.5f pairs %d vs %d" % (D, radii_match, depth_match,
                                                                            rz.last_pair_count(), n_ref))
        assert (R["render_alphas"] > 0.5).float().mean() > 0.02
        torch.testing.assert_close(ra, R["render_alphas"], rtol=0, atol=IMG_ATOL)
        torch.testing.assert_close(rc, R["render_colors"], rtol=0, atol=IMG_ATOL)


def test_fused_multi_camera_equals_single_camera_calls():
    from ubs_b200 import fused, synth

    D, N, W, H, C = 6, 50000, 320, 240, 4
    scene = synth.make_scene(N, D, seed=77).to("cuda")
    cams = synth.make_cameras(C, W, H, seed=5, device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    V = torch.stack([c.viewmat for c in cams])
    K = torch.stack([c.K for c in cams])
    P = torch.stack([c.cam_pos for c in cams])
    bg = torch.rand(C, 3, device="cuda")
    multi = fused.FusedRasterizer(D, N, W, H, n_cams=C)
    rc, ra = multi.forward(rec, V, K, P, None, bg)
    single = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    for c in range(C):
        rc1, ra1 = single.forward(rec, V[c:c + 1], K[c:c + 1], P[c:c + 1], None, bg[c:c + 1])
        assert torch.equal(rc[c], rc1[0]) and torch.equal(ra[c], ra1[0])


def test_fused_capacity_overflow_is_flagged_and_recovers():
    from ubs_b200 import fused, synth

    D, N, W, H = 6, 30000, 320, 240
    scene = synth.make_scene(N, D, seed=3).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=5, device="cuda")[0]
    rec = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1, capacity=1000)
    rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None])
    assert rz.overflowed() and rz.last_pair_count() > 1000
    torch.cuda.synchronize()
    rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None])  # polls the count, grows the buffers
    assert rz.capacity >= rz.last_pair_count()
    full = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    rc, ra = full.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None])
    rc2, ra2 = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None])
    assert torch.equal(rc, rc2) and torch.equal(ra, ra2)


@pytest.mark.parametrize("D,N,W,H", [(6, 50000, 480, 360), (7, 30000, 400, 300)])
def test_fused_backward_matches_reference_chain(D, N, W, H):
    ref = _ref()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import fused, synth

    scene = synth.make_scene(N, D, seed=99 + D).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=4, timestamps=[0.6], device="cuda")[0]
    bg = torch.tensor([0.9, 0.8, 0.7], device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(1, H, W, 1, device="cuda", generator=g) / (H * W)
    ref_grads, R = ref.chain_grads(scene, cam, bg, v_rc, v_ra)

    rec = fused.pack_records(D, *scene.tensors()).requires_grad_(True)
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
    bgl = bg[None].clone().requires_grad_(True)
    rc, ra = fused.render(rec, rz, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bgl)
    torch.autograd.backward((rc, ra), (v_rc, v_ra))
    mine = fused.unpack_records(D, rec.grad)
    names = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")
    for name, a, b in zip(names, mine, ref_grads):
        _assert_grad_close(name, a, b.reshape(a.shape), rtol=3e-3)  # three chained stages of fp32 rounding
    assert (rec.grad[:, fused.record_slices(D)["l_triangle"].stop:] == 0).all()  # padding columns
    v_bg_ref = (v_rc * (1.0 - R["render_alphas"])).sum(dim=(1, 2))
    torch.testing.assert_close(bgl.grad, v_bg_ref, rtol=1e-3, atol=1e-6)


def test_two_threads_two_streams_render_concurrently():
    """SURVEY 8(b) threading row: the training thread renders while a viewer thread calls view() (train.py:95-98,177).
    Two host threads, each on its own CUDA stream with its own FusedRasterizer, render different cameras of one shared
    record buffer at the same time; every frame must be bit-identical to the same camera rendered alone, and a bad call
    in one thread must not leak its error message into the other (thread-local error channel)."""
    import threading

    from ubs_b200 import _lib, fused, synth

    D, N, W, H = 6, 80000, 480, 360
    scene = synth.make_scene(N, D, seed=77).to("cuda")
    cams = synth.make_cameras(6, W, H, seed=5, device="cuda")
    bg = torch.tensor([[0.2, 0.2, 0.2]], device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    rz0 = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    want = []
    for cam in cams:
        rc, ra = rz0.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
        want.append((rc.clone(), ra.clone()))
    torch.cuda.synchronize()

    lib = _lib.load()
    errors, got = [], {}

    def worker(tid):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
                for rep in range(4):
                    for k in range(tid, len(cams), 2):
                        cam = cams[k]
                        rc, ra = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
                        got[(tid, rep, k)] = (rc.clone(), ra.clone())
                    if tid == 1 and rep == 1:
                        # an invalid call on this thread only: stride query with an unsupported dimension
                        assert lib.ubs_record_stride(99) < 0
                        assert lib.ubs_rasterize_fwd(1, 1, None, 0, None, None, None, None, None, None, None, 3, W, H, 7,
                                                     None, None, None, None, None, None) < 0
                        assert b"tile_size" in lib.ubs_last_error()
                stream.synchronize()
            if tid == 0:
                assert b"tile_size" not in lib.ubs_last_error()
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    assert len(got) == 4 * len(cams)
    for (tid, rep, k), (rc, ra) in got.items():
        assert torch.equal(rc, want[k][0]) and torch.equal(ra, want[k][1]), (tid, rep, k)


@pytest.mark.parametrize("D,N,W,H", [(6, 50000, 480, 360), (7, 30000, 400, 300)])
def test_fused_viewmat_gradient_matches_reference_projection_backward(D, N, W, H):
    """viewmats.requires_grad on the fused path (`_wrapper.py:898`): the gradient of the world-to-camera matrix through
    the projection, against the reference's fully_fused_projection_bwd(viewmats_requires_grad=True) inside the reference
    chain (the view direction of the conditioning is detached in both, scene/beta_model.py:675-690)."""
    ref = _ref()
    from test_gpu_backward import _assert_grad_close
    from ubs_b200 import fused, synth

    scene = synth.make_scene(N, D, seed=61 + D).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=4, timestamps=[0.6], device="cuda")[0]
    bg = torch.tensor([0.2, 0.5, 0.7], device="cuda")
    g = torch.Generator(device="cuda").manual_seed(8)
    v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(1, H, W, 1, device="cuda", generator=g) / (H * W)
    keep = {}
    ref_grads, _ = ref.chain_grads(scene, cam, bg, v_rc, v_ra, keep=keep)
    r_view = keep["v_viewmats"]

    rec = fused.pack_records(D, *scene.tensors()).requires_grad_(True)
    vm = cam.viewmat[None].clone().requires_grad_(True)
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
    rc, ra = fused.render(rec, rz, vm, cam.K[None], cam.cam_pos[None], ts, bg[None])
    torch.autograd.backward((rc, ra), (v_rc, v_ra))
    assert vm.grad is not None and vm.grad.shape == (1, 4, 4)
    _assert_grad_close("v_viewmats", vm.grad[:, :3, :], r_view[:, :3, :], rtol=5e-3)
    assert (vm.grad[:, 3, :] == 0).all()
    # the parameter gradients are those of the plain instantiation
    mine = fused.unpack_records(D, rec.grad)
    for name, a, b in zip(("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"), mine, ref_grads):
        _assert_grad_close(name, a, b.reshape(a.shape), rtol=3e-3)


def test_splat_rows_equal_the_separate_arrays_and_both_compositing_entries_agree():
    """The fused projection writes every visible primitive's screen-space record twice: as the separate arrays of the
    reference (radii, means2d, conics, ...) and as one 48-byte row (`splats`) that the compositing kernels gather from.
    The rows must equal the arrays bit for bit, and ubs_rasterize_{fwd,bwd}_splats must give bit-identical images and
    gradients to ubs_rasterize_{fwd,bwd} on the arrays (also with an explicit colour array, the viewer's path)."""
    from ubs_b200 import _lib, fused, ops, synth
    from ubs_b200._lib import check, ptr

    D, N, W, H = 6, 70000, 400, 304
    scene = synth.make_scene(N, D, seed=91).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=8, device="cuda")[0]
    bg = torch.tensor([[0.3, 0.1, 0.6]], device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    rc, ra = rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
    vis = (rz.radii > 0)[0]
    assert int(vis.sum()) > 1000
    sp = rz.splats[0][vis]
    assert torch.equal(sp[:, 0:2], rz.means2d[0][vis]) and torch.equal(sp[:, 2], rz.opacities[0][vis])
    assert torch.equal(sp[:, 3], rz.betas[0][vis]) and torch.equal(sp[:, 4:7], rz.conics[0][vis])
    assert torch.equal(sp[:, 7], rz.depths[0][vis]) and torch.equal(sp[:, 8:11], rz.colors[0][vis])

    # the reference-shaped operator on the separate arrays
    n = rz.last_pair_count()
    rc2, ra2, last2 = ops.rasterize_fwd(rz.means2d, rz.conics, rz.colors, rz.opacities, rz.betas, bg, None, W, H, 16,
                                        rz.offsets, rz.flatten_ids[:n])
    assert torch.equal(rc2, rc) and torch.equal(ra2, ra) and torch.equal(last2, rz.last_ids)

    lib, s = _lib.load(), torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(3)
    v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(1, H, W, 1, device="cuda", generator=g) / (H * W)

    def grads(use_splats, colors):
        out = [torch.zeros_like(t) for t in (rz.means2d, rz.conics, rz.colors, rz.opacities, rz.betas)]
        if use_splats:
            check(lib.ubs_rasterize_bwd_splats(1, N, ptr(rz.n_isects), rz.capacity, ptr(rz.splats), ptr(colors), ptr(bg), None,
                                               3, W, H, 16, ptr(rz.offsets), ptr(rz.flatten_ids), ptr(rz.render_alphas),
                                               ptr(rz.last_ids), ptr(v_rc), ptr(v_ra), *[ptr(t) for t in out], None, s), "bwd_splats")
        else:
            check(lib.ubs_rasterize_bwd(1, N, ptr(rz.n_isects), rz.capacity, ptr(rz.means2d), ptr(rz.conics), ptr(rz.colors),
                                        ptr(rz.opacities), ptr(rz.betas), ptr(bg), None, 3, W, H, 16, ptr(rz.offsets),
                                        ptr(rz.flatten_ids), ptr(rz.render_alphas), ptr(rz.last_ids), ptr(v_rc), ptr(v_ra),
                                        *[ptr(t) for t in out], s), "bwd")
        return out

    # gradients are sums of float atomics: equal up to the order of the additions
    ref = grads(False, None)
    for variant in (grads(True, None), grads(True, rz.colors)):
        for a, b in zip(variant, ref):
            assert (a - b).abs().max().item() <= 1e-6 * b.abs().max().item() + 1e-12
    # forward with an explicit colour array next to the splat rows
    rc3, ra3, last3 = torch.empty_like(rc), torch.empty_like(ra), torch.empty_like(rz.last_ids)
    check(lib.ubs_rasterize_fwd_splats(1, N, ptr(rz.n_isects), rz.capacity, ptr(rz.splats), ptr(rz.colors), ptr(bg), None, 3,
                                       W, H, 16, ptr(rz.offsets), ptr(rz.flatten_ids), ptr(rc3), ptr(ra3), ptr(last3), s),
          "fwd_splats")
    assert torch.equal(rc3, rc) and torch.equal(ra3, ra) and torch.equal(last3, rz.last_ids)


@pytest.mark.parametrize("D,N,W,H,C,aa", [(6, 70000, 400, 304, 1, False), (7, 40000, 333, 250, 2, True),
                                           (6, 3000, 1280, 720, 1, False)])
def test_gradient_rows_equal_the_separate_gradient_arrays(D, N, W, H, C, aa):
    """ubs_rasterize_bwd_rows (lane = pair, 4x4 blocks, vector reductions into one 48-byte row per primitive) against
    ubs_rasterize_bwd_splats (separate arrays): the rows, decoded as include/ubs_b200.h documents, equal the arrays up
    to the order of the float additions; and the projection backward gives the same parameter gradients from either
    (FusedRasterizer(grad_rows=True / False)).  The last case has a few huge splats per tile (long single-pair runs,
    short buckets), the second one partially covered border tiles, two cameras and the antialiased opacities."""
    from ubs_b200 import fused, synth
    from ubs_b200._lib import check, ptr

    scene = synth.make_scene(N, D, seed=17 + D).to("cuda")
    if N < 10000:
        scene.scale.data += 1.5  # large splats: long lists of pairs that cover whole tiles
    cams = synth.make_cameras(C, W, H, seed=6, timestamps=[0.3, 0.8][:C], device="cuda")
    vm, K, cp = (torch.stack([getattr(c, k) for c in cams]) for k in ("viewmat", "K", "cam_pos"))
    ts = torch.tensor([c.timestamp for c in cams], device="cuda") if D == 7 else None
    bg = torch.rand(C, 3, device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    g = torch.Generator(device="cuda").manual_seed(3)
    v_rc = torch.randn(C, H, W, 3, device="cuda", generator=g) / (H * W)
    v_ra = torch.randn(C, H, W, 1, device="cuda", generator=g) / (H * W)
    out = {}
    for rows in (True, False):
        rz = fused.FusedRasterizer(D, N, W, H, n_cams=C, antialiased=aa, grad_rows=rows, capacity=6_000_000)
        rz.forward(rec, vm, K, cp, ts, bg)
        assert not rz.overflowed()
        out[rows] = (rz, rz.backward(rec, vm, K, cp, ts, bg, v_rc, v_ra))
    (rz_r, g_r), (rz_a, g_a) = out[True], out[False]
    assert int((rz_a.radii > 0).sum()) > min(N, 1000) // 2
    r = rz_r.v_rows
    a, b, c = rz_a.conics.unbind(-1)
    decoded = {
        "v_colors": (r[..., 0:3], rz_a.v_colors),
        "v_conics": (torch.stack((r[..., 3], 2 * r[..., 4], r[..., 5]), -1), rz_a.v_conics),
        "v_means2d": (torch.stack((2 * a * r[..., 6] + 2 * b * r[..., 7], 2 * b * r[..., 6] + 2 * c * r[..., 7]), -1),
                      rz_a.v_means2d),
        "v_opacities": (r[..., 8], rz_a.v_opacities),
        "v_betas": (r[..., 9] * math.log(2.0), rz_a.v_betas),
    }
    vis = rz_a.radii > 0
    for name, (x, y) in decoded.items():
        x, y = x[vis].double(), y[vis].double()
        scale = y.abs().max().item()
        assert scale > 0, name
        # sums of up to thousands of float terms in a different order (and a different factoring for v_means2d)
        assert (x - y).abs().max().item() <= 2e-5 * scale, (name, (x - y).abs().max().item(), scale)
        assert ((x - y).abs().sum() / y.abs().sum()).item() < 2e-6, name
    assert (r[..., 10:] == 0).all() and (r[~vis] == 0).all()
    scale = g_a.abs().amax(dim=0).clamp_min(1e-20)
    assert ((g_r - g_a).abs().amax(dim=0) / scale).max().item() < 2e-4


def test_render_only_frames_render_queue_and_host_pipeline_match_plain_forward():
    """screen_space=False skips the arrays only `meta` / backward read; RenderQueue and a two-rasteriser HostPipeline
    run frames on several streams.  All of them must give bit-identical images."""
    from ubs_b200 import fused, synth

    D, N, W, H = 6, 40000, 400, 304
    scene = synth.make_scene(N, D, seed=77).to("cuda")
    cams = synth.make_cameras(5, W, H, seed=5, device="cuda")
    bg = torch.tensor([[0.2, 0.1, 0.4]], device="cuda")
    rec = fused.pack_records(D, *scene.tensors())
    plain = fused.FusedRasterizer(D, N, W, H)
    want = []
    for c in cams:
        rc, ra = plain.forward(rec, c.viewmat[None], c.K[None], c.cam_pos[None], None, bg)
        want.append((rc.clone(), ra.clone()))
    rz = fused.FusedRasterizer(D, N, W, H)
    for c, (rc0, ra0) in zip(cams, want):
        rz.conics.fill_(float("nan"))
        rc, ra = rz.forward(rec, c.viewmat[None], c.K[None], c.cam_pos[None], None, bg, screen_space=False)
        assert torch.equal(rc, rc0) and torch.equal(ra, ra0)
        assert bool(torch.isnan(rz.conics).all()), "render-only frame wrote the separate arrays"
    with pytest.raises(AssertionError):
        rz.backward(rec, cams[0].viewmat[None], cams[0].K[None], cams[0].cam_pos[None], None, bg, rc, ra)
    made = []
    queue = fused.RenderQueue(lambda: made.append(fused.FusedRasterizer(D, N, W, H)) or made[-1], depth=2)
    outs = []
    for c in cams:
        slot, rc, ra = queue.render(rec, c.viewmat[None], c.K[None], c.cam_pos[None], None, bg, screen_space=False)
        queue.wait(slot)
        outs.append((rc.clone(), ra.clone()))
    for (rc, ra), (rc0, ra0) in zip(outs, want):
        assert torch.equal(rc, rc0) and torch.equal(ra, ra0)
    pipe = fused.HostPipeline(made, depth=3)
    rows = torch.empty((len(cams), fused.HostPipeline.CAM_FLOATS)).pin_memory()
    for k, c in enumerate(cams):
        rows[k] = torch.cat([c.viewmat.reshape(-1), c.K.reshape(-1), c.cam_pos, torch.zeros(1, device="cuda")]).cpu()
    for k in range(len(cams)):
        slot = pipe.render_to_host(rec, rows[k], bg)
        host_rc, _ = pipe.wait(slot)
        assert torch.equal(host_rc, want[k][0].cpu())
    pipe.drain()
