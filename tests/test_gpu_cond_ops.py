"""GPU parity of K1-K4 (covariance build, conditioning; forward and backward) against the reference kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-3  # north-star tolerance on gradients (relative)


def _ref():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("reference CUDA oracle not built")
    return ref_cuda


def _rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("D", [6, 7])
@pytest.mark.parametrize("spatial", [False, True])
def test_covariance_fwd_bwd(D, spatial):
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops, synth

    sc = synth.make_scene(20000, D, seed=5 + D).to("cuda")
    scale = torch.nn.functional.softplus(sc.scale)
    ri, rj = ref.tril_rest(D, "cuda")
    rot_ref = C_.l_triangle_to_rotmat_fwd(sc.l_triangle[:, :3].contiguous())
    rot = ops.l_triangle_to_rotmat(sc.l_triangle[:, :3])
    assert torch.equal(rot, rot_ref)
    cov_ref = C_.rot_scale_l_triangle_to_covar_fwd(rot_ref, scale, sc.l_triangle, ri, rj, spatial)
    cov = ops.rot_scale_l_triangle_to_covar(rot, scale, sc.l_triangle, ri, rj, spatial)
    assert cov.shape == cov_ref.shape
    assert torch.equal(cov, cov_ref), "covariance not bit-exact: max diff %g" % (cov - cov_ref).abs().max().item()
    g = torch.randn_like(cov_ref)
    vr_ref, vs_ref, vl_ref = C_.rot_scale_l_triangle_to_covar_bwd(rot_ref, scale, sc.l_triangle, ri, rj, spatial, g)
    rot_l = rot.detach().requires_grad_(True)
    scale_l = scale.detach().requires_grad_(True)
    lt_l = sc.l_triangle.detach().requires_grad_(True)
    ops.rot_scale_l_triangle_to_covar(rot_l, scale_l, lt_l, ri, rj, spatial).backward(g)
    assert _rel_err(rot_l.grad, vr_ref) < GRAD_RTOL
    assert _rel_err(scale_l.grad, vs_ref) < GRAD_RTOL
    assert _rel_err(lt_l.grad, vl_ref) < GRAD_RTOL
    # K1 backward
    gR = torch.randn_like(rot_ref)
    l3 = sc.l_triangle[:, :3].detach().clone().requires_grad_(True)
    ops.l_triangle_to_rotmat(l3).backward(gR)
    assert torch.equal(l3.grad, C_.l_triangle_to_rotmat_bwd(l3.detach(), gR))


@pytest.mark.parametrize("D", [6, 7])
def test_conditioning_fwd_bwd(D):
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops, synth

    sc = synth.make_scene(30000, D, seed=77 + D).to("cuda")
    cam = synth.make_cameras(1, 640, 480, seed=3)[0]
    cam.cam_pos = cam.cam_pos.cuda()
    cam.timestamp = 0.37
    scale, opacity, beta, mean = ref.activations(sc)
    ri, rj = ref.tril_rest(D, "cuda")
    rot = C_.l_triangle_to_rotmat_fwd(sc.l_triangle[:, :3].contiguous())
    covar = C_.rot_scale_l_triangle_to_covar_fwd(rot, scale, sc.l_triangle, ri, rj, False)
    q = ref.query_for(sc, cam).contiguous()
    b = beta[:, 1:].contiguous()
    m_ref, v_ref, o_ref = C_.cond_mean_convariance_opacity_fwd(mean, covar, opacity, b, q)
    m, v, o = ops.cond_mean_convariance_opacity(mean, covar, opacity, b, q)
    bitexact = (m == m_ref).float().mean().item()
    torch.testing.assert_close(m, m_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(v, v_ref, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(o, o_ref, rtol=1e-4, atol=1e-7)
    assert bitexact > 0.99, "conditional mean bit-exact rate %.4f" % bitexact
    gm, gv, go = torch.randn_like(m_ref), torch.randn_like(v_ref), torch.randn_like(o_ref)
    r_means, r_covars, r_opac, r_betas = C_.cond_mean_convariance_opacity_bwd(mean, covar, opacity, b, q, gm, gv, go)
    ml, cl, ol, bl = [t.detach().clone().requires_grad_(True) for t in (mean, covar, opacity, b)]
    ql = q.detach().clone().requires_grad_(True)
    outs = ops.cond_mean_convariance_opacity(ml, cl, ol, bl, ql)
    torch.autograd.backward(outs, (gm, gv, go))
    assert ql.grad is None  # the reference returns None for query (cuda/_wrapper.py:611)
    for name, mine, theirs in (("means", ml.grad, r_means), ("covars", cl.grad, r_covars), ("opac", ol.grad, r_opac),
                               ("betas", bl.grad, r_betas)):
        # per-primitive relative error, robust to a few ill-conditioned primitives
        num = (mine - theirs).flatten(1).abs().amax(1)
        den = theirs.flatten(1).abs().amax(1).clamp_min(1e-12)
        bad = (num / den > GRAD_RTOL).float().mean().item()
        assert bad < 2e-3, "%s: %.4f%% of primitives exceed rel %g" % (name, 100 * bad, GRAD_RTOL)
        assert _rel_err(mine, theirs) < GRAD_RTOL, name


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_companion_operators_accept_half_and_bfloat16(dtype):
    """The reference dispatches K1-K4 over half / bfloat16 too (cond_mean_convariance_opacity_fwd.cu:330,
    rot_scale_l_triangle_to_covar_fwd.cu:216).  Here reduced-precision inputs are widened, computed by the float32
    kernels and narrowed back: outputs and gradients arrive in the input type and equal the float32 results of the same
    (already rounded) inputs after one rounding; K1 / K2, which have no cancellation, also agree with the reference's own
    reduced-precision kernels to a few units of the type's precision."""
    ref = _ref()
    C_ = ref.load()
    from ubs_b200 import ops, synth

    D, N = 6, 5000
    sc = synth.make_scene(N, D, seed=12).to("cuda")
    lt = (0.3 * sc.l_triangle).to(dtype)
    scale = torch.nn.functional.softplus(sc.scale).to(dtype)
    ri, rj = ref.tril_rest(D, "cuda")
    lt_l, scale_l = lt.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    rot = ops.l_triangle_to_rotmat(lt_l[:, :3])
    cov = ops.rot_scale_l_triangle_to_covar(rot, scale_l, lt_l, ri, rj, False)
    assert rot.dtype == dtype and cov.dtype == dtype and cov.shape == (N, D, D)
    # against float32 on the same inputs
    rot32 = ops.l_triangle_to_rotmat(lt.float()[:, :3].contiguous())
    cov32 = ops.rot_scale_l_triangle_to_covar(rot32, scale.float(), lt.float(), ri, rj, False)
    assert torch.equal(rot, rot32.to(dtype))
    eps = torch.finfo(dtype).eps
    assert ((cov.float() - cov32).abs() <= 4 * eps * cov32.abs().amax(dim=(1, 2), keepdim=True)).all()
    # against the reference's reduced-precision instantiations
    r_rot = C_.l_triangle_to_rotmat_fwd(lt[:, :3].contiguous())
    assert r_rot.dtype == dtype and torch.equal(r_rot, rot.detach())
    r_cov = C_.rot_scale_l_triangle_to_covar_fwd(r_rot, scale, lt, ri, rj, False)
    tol = 16 * eps * cov32.abs().amax(dim=(1, 2), keepdim=True)
    assert ((r_cov.float() - cov.detach().float()).abs() <= tol).all()
    # conditioning + gradients in the input type
    mean = torch.cat([sc.xyz, sc.mean], dim=-1).to(dtype)
    o = torch.sigmoid(sc.opacity).to(dtype)
    b = (4.0 * torch.exp(sc.beta))[:, 1:].contiguous().to(dtype)
    q = torch.nn.functional.normalize(torch.randn(N, D - 3, device="cuda"), dim=-1).to(dtype)
    m3, v3, o3 = ops.cond_mean_convariance_opacity(mean, cov, o, b, q)
    assert m3.dtype == v3.dtype == o3.dtype == dtype and v3.shape == (N, 3, 3)
    m32, v32, o32 = ops.cond_mean_convariance_opacity(mean.float(), cov.detach().float(), o.float(), b.float(), q.float())
    assert torch.equal(m3.detach(), m32.to(dtype)) and torch.equal(o3.detach(), o32.to(dtype))
    (m3.float().sum() + v3.float().sum() + o3.float().sum()).backward()
    assert lt_l.grad is not None and lt_l.grad.dtype == dtype and scale_l.grad.dtype == dtype
    assert torch.isfinite(lt_l.grad.float()).all() and torch.isfinite(scale_l.grad.float()).all()
    with pytest.raises(RuntimeError):
        ops.l_triangle_to_rotmat(lt.double()[:, :3].contiguous())
