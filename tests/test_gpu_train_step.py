"""GPU parity of the train-step kernels (csrc/loss.cu, csrc/optim.cu) through the C-ABI against the CPU oracle
(oracle/train_oracle.py) and the committed reference fixture tests/golden/loss_ssim.npz.

Tolerances (FP32): loss values 2e-6 absolute; image gradient 2e-5 of its max-abs (separable 11+11-tap window vs the
reference's dense 121-tap conv2d: different summation order); Adam 2e-6 relative after several steps; relocation
1e-6 on the rescaled logits, everything else bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_ssim.npz")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_loss_kernel_matches_reference_fixture(tag):
    from ubs_b200 import training

    z = np.load(GOLD)
    img = torch.from_numpy(z[f"{tag}_img"]).cuda()[None]  # [1,ch,H,W]
    gt = torch.from_numpy(z[f"{tag}_gt"]).cuda()[None]
    out, v = training.l1_ssim_loss_fwd_bwd(img, gt, 0.2, 1.0, "NCHW", "NCHW")
    l1, ssim, loss = out.tolist()
    assert abs(l1 - float(z[f"{tag}_l1"])) < 2e-6
    assert abs(ssim - float(z[f"{tag}_ssim"])) < 2e-6
    assert abs(loss - float(z[f"{tag}_loss"])) < 2e-6
    ref = torch.from_numpy(z[f"{tag}_grad"])
    err = (v[0].cpu() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), err


@pytest.mark.parametrize("C,H,W", [(1, 200, 333), (2, 65, 31), (1, 1080, 1920)])
def test_loss_kernel_nhwc_render_layout_vs_oracle(C, H, W):
    """The layout the train step uses: rendered [C,H,W,3] image against a [C,3,H,W] ground truth, gradient scaled."""
    from oracle import train_oracle as T
    from ubs_b200 import training

    g = torch.Generator().manual_seed(H * 7 + W)
    gt = torch.rand((C, 3, H, W), generator=g)
    img = (gt + 0.1 * torch.randn((C, 3, H, W), generator=g)).clamp(0, 1)
    img_o = img.clone().requires_grad_(True)
    loss_o = T.photometric_loss(img_o, gt, 0.35)
    (0.5 * loss_o).backward()
    rc = img.permute(0, 2, 3, 1).contiguous().cuda()
    out, v = training.l1_ssim_loss_fwd_bwd(rc, gt.cuda(), 0.35, 0.5, "NHWC", "NCHW")
    assert v.shape == rc.shape
    assert abs(out[2].item() - loss_o.item()) < 2e-6
    ref = img_o.grad.permute(0, 2, 3, 1)
    err = (v.cpu() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), err
    # evaluate-only call (no gradient, small workspace path) gives the same numbers
    out2, v2 = training.l1_ssim_loss_fwd_bwd(rc, gt.cuda(), 0.35, 1.0, "NHWC", "NCHW", want_grad=False)
    assert v2 is None and torch.equal(out2, out)


def test_loss_autograd_function_matches_oracle():
    from oracle import train_oracle as T
    from ubs_b200 import training

    g = torch.Generator().manual_seed(5)
    gt = torch.rand((3, 90, 70), generator=g)
    img = torch.rand((3, 90, 70), generator=g)
    a = img.clone().requires_grad_(True)
    T.photometric_loss(a, gt).backward()
    b = img.cuda().requires_grad_(True)
    loss = training.l1_ssim_loss(b, gt.cuda())
    (3.0 * loss).backward()
    assert (b.grad.cpu() / 3.0 - a.grad).abs().max().item() <= 2e-5 * a.grad.abs().max().item()


@pytest.mark.parametrize("D,N", [(6, 5000), (7, 3001)])
def test_packed_adam_matches_torch_adam(D, N):
    """Five steps with a moving xyz learning rate and the regularisers on, against torch.optim.Adam on the seven
    separate tensors (scene/beta_model.py:239-268) with the regularisers differentiated by autograd."""
    from oracle import train_oracle as T
    from ubs_b200 import fused, synth, training

    scene = synth.make_scene(N, D, seed=77)
    params = [t.clone().requires_grad_(True) for t in scene.tensors()]
    lr = dict(training.DEFAULT_LR)
    opt = T.make_adam(params, lr)
    rec = fused.pack_records(D, *[t.cuda() for t in scene.tensors()])
    adam = training.PackedAdam(D, N, lr)
    g = torch.Generator().manual_seed(3)
    for it in range(5):
        xyz_lr = training.expon_lr(it * 3000, 1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30000)
        opt.param_groups[0]["lr"] = xyz_lr
        adam.set_lr("xyz", xyz_lr)
        grads = [torch.randn(p.shape, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
                 for p in params]
        for gr in grads:
            gr[torch.rand(gr.shape[0], generator=g) < 0.4] = 0.0  # culled primitives have zero gradient rows
        opt.zero_grad()
        T.regularisers(params[3], params[5], 0.01, 0.02).backward()
        for p, gr in zip(params, grads):
            p.grad = gr.clone() if p.grad is None else p.grad + gr.reshape(p.grad.shape)
        opt.step()
        grec = fused.pack_records(D, *[gr.cuda() for gr in grads])
        adam.step(rec, grec, opacity_reg=0.01, scale_reg=0.02)
    for name, mine, p in zip(T.GROUPS, fused.unpack_records(D, rec.cpu()), params):
        ref = p.detach().reshape(mine.shape)
        # 2e-6 relative, plus 1e-5 of the distance five steps can move a parameter (Adam's update is lr * m / sqrt(v):
        # a rounding-level difference in a tiny gradient is amplified to a rounding-level fraction of lr)
        tol = 2e-6 * ref.abs() + 1e-5 * 5 * lr[name]
        assert bool(((mine - ref).abs() <= tol).all()), (name, (mine - ref).abs().max().item())
    sl = fused.record_slices(D)
    st = opt.state[params[5]]
    assert torch.allclose(adam.exp_avg[:, sl["scale"]].cpu(), st["exp_avg"], rtol=2e-6, atol=1e-12)
    assert torch.allclose(adam.exp_avg_sq[:, sl["scale"]].cpu(), st["exp_avg_sq"], rtol=2e-6, atol=1e-20)
    # padding columns never move
    pad = rec[:, fused.record_slices(D)["l_triangle"].stop:]
    assert pad.numel() == 0 or float(pad.abs().max()) == 0.0


@pytest.mark.parametrize("D", [6, 7])
def test_mcmc_relocate_and_add_match_oracle(D):
    from oracle import train_oracle as T
    from ubs_b200 import fused, synth, training

    N = 4000
    scene = synth.make_scene(N, D, seed=5)
    params = [t.clone() for t in scene.tensors()]
    params[3] = params[3].reshape(N, 1)
    g = torch.Generator().manual_seed(9)
    params[3][torch.rand(N, generator=g) < 0.1] = -8.0  # dead primitives: sigmoid(-8) < 0.005
    rec = fused.pack_records(D, *[p.cuda() for p in params])
    adam = training.PackedAdam(D, N)
    adam.exp_avg.normal_()
    adam.exp_avg_sq.uniform_()
    sl = fused.record_slices(D)
    names = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")
    moments = [(adam.exp_avg[:, sl[n]].cpu().clone(), adam.exp_avg_sq[:, sl[n]].cpu().clone()) for n in names]

    dead_mask = torch.sigmoid(params[3][:, 0]) <= 0.005
    dead = dead_mask.nonzero(as_tuple=True)[0]
    alive = (~dead_mask).nonzero(as_tuple=True)[0]
    src = training.sample_alive(torch.sigmoid(params[3][alive, 0]), dead.numel(), alive, generator=g)
    assert dead.numel() > 100 and torch.bincount(src).max() >= 2
    T.relocate(params, moments, dead, src)
    training.mcmc_relocate(rec, D, dead.cuda(), src.cuda(), adam)

    def compare(rec_gpu, params_cpu, moments_cpu):
        for n, mine, p in zip(names, fused.unpack_records(D, rec_gpu.cpu()), params_cpu):
            if n == "opacity":
                assert (mine - p).abs().max().item() < 1e-5, n
            else:
                assert torch.equal(mine, p.reshape(mine.shape)), n
        for n, (m, v) in zip(names, moments_cpu):
            assert torch.equal(adam.exp_avg[:, sl[n]].cpu(), m), n
            assert torch.equal(adam.exp_avg_sq[:, sl[n]].cpu(), v), n

    compare(rec, params, moments)
    add_idx = training.sample_alive(torch.sigmoid(params[3][:, 0]), 137, generator=g)
    grown, gm = T.add_new(params, moments, add_idx)
    rec2 = training.mcmc_add(rec, D, add_idx.cuda(), adam)
    assert rec2.shape[0] == N + 137 and adam.exp_avg.shape[0] == N + 137
    compare(rec2, grown, gm)


@pytest.mark.parametrize("D", [6, 7])
def test_mcmc_relocate_and_add_match_reference_fixture(D):
    """ubs_mcmc_relocate (training.mcmc_relocate / mcmc_add) against the outputs of the reference's own relocate_gs /
    add_new_gs (tests/golden/relocate_D*.npz, generated by tests/golden/make_golden_relocate.py from the reference
    source): copied rows bit-exact, rescaled opacity logits within 1e-5, Adam moments bit-exact."""
    import os

    import numpy as np

    from ubs_b200 import fused, training

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"relocate_D{D}.npz"))
    names = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")

    def pack(tag, suffix=""):
        return fused.pack_records(D, *[torch.from_numpy(z[f"{tag}_{n}{suffix}"]).cuda() for n in names])

    rec = pack("in")
    N = rec.shape[0]
    adam = training.PackedAdam(D, N)
    adam.exp_avg.copy_(pack("in", "_m"))
    adam.exp_avg_sq.copy_(pack("in", "_v"))
    dead = torch.from_numpy(z["dead_mask"]).nonzero(as_tuple=True)[0].cuda()
    training.mcmc_relocate(rec, D, dead, torch.from_numpy(z["reinit_idx"]).cuda(), adam)

    def compare(rec_gpu, tag):
        sl = fused.record_slices(D)
        want, want_m, want_v = pack(tag), pack(tag, "_m"), pack(tag, "_v")
        assert rec_gpu.shape == want.shape
        for n in names:
            a, b = rec_gpu[:, sl[n]], want[:, sl[n]]
            if n == "opacity":
                assert (a - b).abs().max().item() < 1e-5, n
            else:
                assert torch.equal(a, b), n
        assert torch.equal(adam.exp_avg, want_m) and torch.equal(adam.exp_avg_sq, want_v)

    compare(rec, "relocated")
    rec2 = training.mcmc_add(rec, D, torch.from_numpy(z["add_idx"]).cuda(), adam)
    assert rec2.shape[0] == N + int(z["n_added"])
    compare(rec2, "grown")


@pytest.mark.parametrize("D", [6, 7])
def test_sgld_noise_matches_oracle(D):
    """ubs_sgld_noise against the restated train.py:156-163 on the same N(0,1) draw (FP32 tolerance: 1e-5 relative to
    the largest displacement of the row + 1e-7 absolute; every other record column untouched bit for bit)."""
    from oracle import train_oracle as T
    from ubs_b200 import fused, synth, training

    N = 5000
    scene = synth.make_scene(N, D, seed=17)
    params = [t.clone() for t in scene.tensors()]
    params[3] = params[3].reshape(N, 1)
    g = torch.Generator().manual_seed(4)
    params[3][torch.rand(N, generator=g) < 0.5] -= 6.0  # plenty of nearly transparent primitives: (1 - o)^100 ~ 1
    noise = torch.randn(N, 3, generator=g)
    noise_lr, xyz_lr = 5e5, 1.6e-4 * 0.37
    want = T.sgld_noise(params, noise, noise_lr, xyz_lr)
    rec = fused.pack_records(D, *[p.cuda() for p in params])
    before = rec.clone()
    out = training.sgld_noise(rec, D, noise_lr, xyz_lr, noise=noise.cuda())
    assert torch.equal(out.cpu(), noise)
    got = rec[:, :3].cpu()
    moved = (want - params[0]).abs()
    assert moved.max().item() > 1e-3  # the test exercises a visible displacement
    tol = 1e-5 * moved.max(dim=1, keepdim=True).values + 1e-7 + 2e-7 * params[0].abs()
    assert bool(((got - want).abs() <= tol).all()), (got - want).abs().max().item()
    assert torch.equal(rec[:, 3:], before[:, 3:])
    # default draw: torch's generator on the device defines the noise
    gd = torch.Generator(device="cuda").manual_seed(11)
    rec2 = before.clone()
    n2 = training.sgld_noise(rec2, D, noise_lr, xyz_lr, generator=gd)
    gd.manual_seed(11)
    assert torch.equal(n2, torch.randn((N, 3), device="cuda", generator=gd))


@pytest.mark.parametrize("D", [6, 7])
def test_sgld_noise_matches_reference_fixture(D):
    """ubs_sgld_noise against tests/golden/sgld_D*.npz (the reference's own train.py:156-163 statements, executed by
    tests/golden/make_golden_sgld.py): 1e-5 of the displacement + FP32 rounding of the position."""
    import os

    import numpy as np

    from ubs_b200 import fused, training

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"sgld_D{D}.npz"))
    N = z["xyz"].shape[0]
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    rec = fused.pack_records(D, t(z["xyz"]), torch.zeros(N, D - 3, device="cuda"), torch.zeros(N, 3, device="cuda"),
                             t(z["opacity"]), torch.zeros(N, D - 2, device="cuda"), t(z["scale"]), t(z["l_triangle"]))
    training.sgld_noise(rec, D, float(z["noise_lr"]), float(z["xyz_lr"]), noise=t(z["noise"]))
    want, x0 = torch.from_numpy(z["xyz_out"]), torch.from_numpy(z["xyz"])
    moved = (want - x0).abs()
    got = rec[:, :3].cpu()
    assert bool(((got - want).abs() <= 1e-5 * moved.max(dim=1, keepdim=True).values + 2e-7 * x0.abs() + 1e-12).all()), \
        (got - want).abs().max().item()


def test_train_step_reduces_loss_and_matches_manual_composition():
    """TrainStep.step == forward, loss kernel, backward, Adam composed by hand; and it learns."""
    from ubs_b200 import fused, synth, training

    D, N, W, H = 6, 30000, 320, 240
    scene = synth.make_scene(N, D, seed=21).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=2, device="cuda")[0]
    bg = torch.ones(1, 3, device="cuda")
    rec_gt = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
    gt = rz.forward(rec_gt, *args)[0].clone().permute(0, 3, 1, 2).contiguous()  # [1,3,H,W]
    rec = rec_gt.clone()
    sl = fused.record_slices(D)
    rec[:, sl["rgb"]] += 0.2 * torch.randn_like(rec[:, sl["rgb"]])

    # one step by hand
    rec_a = rec.clone()
    adam_a = training.PackedAdam(D, N)
    rc, _ = rz.forward(rec_a, *args)
    out, v_rc = training.l1_ssim_loss_fwd_bwd(rc, gt, 0.2, 1.0, "NHWC", "NCHW")
    v_rec = rz.backward(rec_a, *args, v_rc, torch.zeros_like(rz.render_alphas))
    adam_a.step(rec_a, v_rec, 0.01, 0.01)
    # the same step through TrainStep
    rec_b = rec.clone()
    ts = training.TrainStep(rz, training.PackedAdam(D, N), fuse_adam=False)
    loss0 = ts.step(rec_b, *args, gt, opacity_reg=0.01, scale_reg=0.01).clone()
    assert torch.equal(loss0, out)
    # the compositing backward accumulates with atomics: a first Adam step is sign-like (lr * g / (|g| + 1e-15)), so a
    # gradient that cancels to rounding noise may flip; everything else must agree
    assert ((rec_a - rec_b).abs() > 1e-7).float().mean().item() < 1e-3
    losses = [loss0[2].item()]
    for _ in range(30):
        losses.append(ts.step(rec_b, *args, gt)[2].item())
    assert losses[-1] < 0.6 * losses[0], losses


@pytest.mark.parametrize("D,N", [(6, 30000), (7, 20011)])
def test_fused_projection_backward_adam_is_bit_identical_to_backward_then_adam(D, N):
    """ubs_fused_project_bwd_adam == ubs_fused_project_bwd followed by ubs_adam_step, on the same screen-space
    gradients (taken from one compositing backward, so the comparison is deterministic), over three steps."""
    import ctypes

    from ubs_b200 import fused, synth, training
    from ubs_b200._lib import check, ptr

    W, H = 320, 240
    scene = synth.make_scene(N, D, seed=40 + D).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=3, timestamps=[0.4], device="cuda")[0]
    bg = torch.zeros(1, 3, device="cuda")
    ts = torch.tensor([0.4], device="cuda") if D == 7 else None
    args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg)
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    rec_a = fused.pack_records(D, *scene.tensors())
    rec_b = rec_a.clone()
    adam_a, adam_b = training.PackedAdam(D, N), training.PackedAdam(D, N)
    g = torch.Generator(device="cuda").manual_seed(1)
    s = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        assert torch.equal(rec_a, rec_b)
        rz.forward(rec_a, *args)
        v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (W * H)
        v_rec = rz.backward(rec_a, *args, v_rc, torch.zeros(1, H, W, 1, device="cuda"))  # leaves rz.v_* filled
        assert int((v_rec.abs().sum(1) > 0).sum()) > N // 20
        adam_a.step(rec_a, v_rec, 0.01, 0.02)
        adam_b.step_count += 1
        cols = (ctypes.c_double * adam_b.stride)(*adam_b.lr_columns())
        check(rz.lib.ubs_fused_project_bwd_adam(
            1, N, D, ptr(rec_b), ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(ts), W, H, rz.eps2d, 0, ptr(rz.radii),
            ptr(rz.conics), *rz.grad_args(), ptr(adam_b.exp_avg), ptr(adam_b.exp_avg_sq), ctypes.cast(cols, ctypes.c_void_p), 0.9,
            0.999, 1e-15, adam_b.step_count, 0.01, 0.02, ptr(rz.status), s), "ubs_fused_project_bwd_adam")
        assert torch.equal(rec_a, rec_b), it
        assert torch.equal(adam_a.exp_avg, adam_b.exp_avg) and torch.equal(adam_a.exp_avg_sq, adam_b.exp_avg_sq), it


def test_train_step_fused_and_unfused_agree():
    from ubs_b200 import fused, synth, training

    D, N, W, H = 6, 30000, 320, 240
    scene = synth.make_scene(N, D, seed=22).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=2, device="cuda")[0]
    bg = torch.ones(1, 3, device="cuda")
    rec0 = fused.pack_records(D, *scene.tensors())
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
    gt = torch.rand(1, 3, H, W, device="cuda")
    out = []
    for fuse in (True, False):
        rec = rec0.clone()
        ts = training.TrainStep(rz, training.PackedAdam(D, N), fuse_adam=fuse)
        for _ in range(3):
            loss = ts.step(rec, *args, gt, opacity_reg=0.01, scale_reg=0.01)[2].item()
        out.append((rec, loss))
    assert abs(out[0][1] - out[1][1]) < 1e-5
    # atomics in the compositing backward: rounding-level gradient noise, amplified only where a gradient ~ 0
    assert ((out[0][0] - out[1][0]).abs() > 1e-6).float().mean().item() < 1e-3


def test_chunked_backward_and_row_range_adam_are_bit_identical_to_whole_buffer_calls():
    """project_backward_rows over row chunks == one whole-buffer call, and Adam applied chunk by chunk (rows=...)
    == one whole-buffer step: the pieces the multi-GPU pipeline (parallel.pipelined_backward) is made of."""
    from ubs_b200 import fused, parallel, synth, training

    D, N, W, H = 6, 25003, 320, 240
    scene = synth.make_scene(N, D, seed=61).to("cuda")
    cam = synth.make_cameras(1, W, H, seed=5, device="cuda")[0]
    bg = torch.zeros(1, 3, device="cuda")
    args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg)
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1)
    rec = fused.pack_records(D, *scene.tensors())
    rz.forward(rec, *args)
    v_rc = torch.randn(1, H, W, 3, device="cuda") / (W * H)
    whole = rz.backward(rec, *args, v_rc, torch.zeros(1, H, W, 1, device="cuda")).clone()  # leaves rz.v_* filled
    chunks = parallel.row_chunks(N, 5)
    assert len(chunks) == 5
    pieces = torch.full_like(whole, float("nan"))
    for b, c in chunks:
        rz.project_backward_rows(rec, *args[:4], pieces, b, c)
    assert torch.equal(pieces, whole)
    # world = 1 goes through the same code path as world > 1, minus the collective
    seen = []
    piped = parallel.pipelined_backward(rz, rec, *args, v_rc, torch.zeros(1, H, W, 1, device="cuda"),
                                        torch.empty_like(whole), 1, None, 5, lambda b, c: seen.append((b, c)))
    assert seen == [(0, N)]
    assert ((piped - whole).abs() > 1e-6 * whole.abs().max()).float().mean().item() < 1e-3  # atomics in compositing

    rec_a, rec_b = rec.clone(), rec.clone()
    adam_a, adam_b = training.PackedAdam(D, N), training.PackedAdam(D, N)
    for it in range(2):
        adam_a.step(rec_a, whole, 0.01, 0.02)
        for k, (b, c) in enumerate(chunks):
            adam_b.step(rec_b, whole, 0.01, 0.02, rows=(b, c), advance=(k == 0))
        assert adam_a.step_count == adam_b.step_count == it + 1
        assert torch.equal(rec_a, rec_b) and torch.equal(adam_a.exp_avg_sq, adam_b.exp_avg_sq)


@pytest.mark.parametrize("world,D,N", [(2, 6, 10007), (3, 7, 6500), (1, 6, 4000)])
def test_sharded_step_kernels_equal_allreduce_then_adam_bitwise(world, D, N):
    """The multi-GPU sharded step (projection backward scattering gradient tiles into the owners' staging slots,
    owner-side reduce + Adam + parameter gather) with all ranks simulated in ONE process on ordinary tensors -- the
    kernels only see device addresses -- against: per-rank gradient records summed in rank order, then the
    whole-buffer Adam.  Bit-identical parameters on every rank, moments equal shard by shard."""
    from ubs_b200 import fused, parallel, synth, training

    W, H = 256, 192
    scene = synth.make_scene(N, D, seed=70 + world).to("cuda")
    cams = synth.make_cameras(world, W, H, seed=6, timestamps=[0.1 * (k + 1) for k in range(world)], device="cuda")
    bg = torch.zeros(1, 3, device="cuda")
    rec0 = fused.pack_records(D, *scene.tensors())
    states = parallel.ShardedState.create_local_group(D, N, world)
    for st in states:
        st.records.copy_(rec0)
        assert st.shard_rows % 128 == 0 and st.shard_rows * world >= N
    rzs = [fused.FusedRasterizer(D, N, W, H, n_cams=1) for _ in range(world)]
    ref_rec = rec0.clone()
    ref_adam = training.PackedAdam(D, N)
    adams = [training.PackedAdam(D, N, allocate_moments=False) for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(2)
    for it in range(3):
        total = torch.zeros_like(ref_rec)
        for r, (st, rz, cam) in enumerate(zip(states, rzs, cams)):
            ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
            args = (cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts)
            rz.forward(st.records, *args, bg)
            v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (W * H)
            rz.composite_backward(bg, v_rc, torch.zeros(1, H, W, 1, device="cuda"))
            parallel.sharded_backward_scatter(rz, st, *args)
            whole = torch.empty_like(ref_rec)
            rz.project_backward_rows(st.records, *args, whole, 0, N)  # same screen-space gradients, plain layout
            total = total + whole
        # every slot of every staging buffer now holds that rank's rows
        for r, st in enumerate(states):
            parallel.sharded_reduce_adam_gather(st, adams[r], 0.01, 0.02)
        ref_adam.step(ref_rec, total, 0.01, 0.02)
        for r, st in enumerate(states):
            assert torch.equal(st.records, ref_rec), (it, r)
            b, n = st.my_rows()
            assert torch.equal(st.exp_avg[:n], ref_adam.exp_avg[b:b + n]), (it, r)
            assert torch.equal(st.exp_avg_sq[:n], ref_adam.exp_avg_sq[b:b + n]), (it, r)
            assert float(st.exp_avg[n:].abs().sum()) == 0.0  # padding rows of the last shard never move


@pytest.mark.parametrize("world,D,N", [(4, 6, 30011), (2, 7, 20000), (8, 6, 9000), (3, 7, 6500), (1, 6, 4000)])
def test_sharded_pull_step_equals_the_multi_camera_fused_update(world, D, N):
    """Pull form of the sharded step (every rank leaves its view's 48-byte gradient rows in a buffer its peers can read;
    the owner of a shard fetches all views' rows of its primitives, runs projection backward + Adam on them and stores
    the new rows into every rank's records), all ranks simulated in ONE process on ordinary tensors, against the
    single-GPU ubs_fused_project_bwd_adam over the same `world` views as one batch, fed the same gradient rows.
    The owner recomputes each view's conic instead of reading the forward's array and tells visibility from the row
    being non-zero; parameters and moments must agree to rounding on every rank."""
    import ctypes

    from ubs_b200 import fused, parallel, synth, training
    from ubs_b200._lib import check, ptr

    W, H = 256, 192
    scene = synth.make_scene(N, D, seed=50 + world).to("cuda")
    cams = synth.make_cameras(world, W, H, seed=9, timestamps=[0.1 * (k + 1) for k in range(world)], device="cuda")
    vm, K, cp = (torch.stack([getattr(c, k) for c in cams]).contiguous() for k in ("viewmat", "K", "cam_pos"))
    ts = torch.tensor([c.timestamp for c in cams], device="cuda") if D == 7 else None
    bg = torch.rand(world, 3, device="cuda")
    rec0 = fused.pack_records(D, *scene.tensors())
    states = parallel.ShardedState.create_local_group(D, N, world)
    rzs = [fused.FusedRasterizer(D, N, W, H, n_cams=1) for _ in range(world)]
    for st, rz in zip(states, rzs):
        st.records.copy_(rec0)
        assert st.exchange == "pull"
        st.attach(rz)
    rz_ref = fused.FusedRasterizer(D, N, W, H, n_cams=world)
    ref_rec, ref_adam = rec0.clone(), training.PackedAdam(D, N)
    adams = [training.PackedAdam(D, N, allocate_moments=False) for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(4)
    s = torch.cuda.current_stream().cuda_stream
    bitwise = True
    for it in range(3):
        for r, (st, rz) in enumerate(zip(states, rzs)):
            one = (vm[r:r + 1], K[r:r + 1], cp[r:r + 1], None if ts is None else ts[r:r + 1])
            rz.forward(st.records, *one, bg[r:r + 1])
            v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / (W * H)
            rz.composite_backward(bg[r:r + 1], v_rc, torch.zeros(1, H, W, 1, device="cuda"))
            assert rz.v_rows.data_ptr() == st.rows.data_ptr() and float(st.rows.abs().sum()) > 0
        # reference: the world views as ONE batch on one GPU, same rows
        rz_ref.forward(ref_rec, vm, K, cp, ts, bg)
        rows_all = torch.cat([st.rows for st in states]).contiguous()
        ref_adam.step_count += 1
        cols = (ctypes.c_double * ref_adam.stride)(*ref_adam.lr_columns())
        check(rz_ref.lib.ubs_fused_project_bwd_adam(
            world, N, D, ptr(ref_rec), ptr(vm), ptr(K), ptr(cp), ptr(ts), W, H, rz_ref.eps2d, 0, ptr(rz_ref.radii),
            ptr(rz_ref.conics), ptr(rows_all), 1, ptr(ref_adam.exp_avg), ptr(ref_adam.exp_avg_sq),
            ctypes.cast(cols, ctypes.c_void_p), 0.9, 0.999, 1e-15, ref_adam.step_count, 0.01, 0.02, None, s),
            "ubs_fused_project_bwd_adam")
        for r, (st, rz) in enumerate(zip(states, rzs)):
            parallel.sharded_pull_update(rz, st, adams[r], vm, K, cp, ts, 0.01, 0.02)
        for r, st in enumerate(states):
            bitwise &= torch.equal(st.records, ref_rec)
            assert torch.equal(st.records, states[0].records), (it, r)  # replicas stay identical
            # Adam's first steps move every touched parameter by ~lr whatever the gradient's size: compare in units of
            # the largest step taken
            step = (ref_rec - rec0).abs().max().item()
            assert (st.records - ref_rec).abs().max().item() <= 2e-3 * step, (it, r)
            b, n = st.my_rows()
            scale = ref_adam.exp_avg.abs().max().item()
            assert (st.exp_avg[:n] - ref_adam.exp_avg[b:b + n]).abs().max().item() <= 1e-5 * scale, (it, r)
            assert float(st.exp_avg[n:].abs().sum()) == 0.0  # padding rows of the last shard never move
    print("pull form bit-identical to the batched single-GPU update:", bitwise)
