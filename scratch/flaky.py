import sys, math, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import ref_cuda as ref
import ubs_b200
from test_gpu_forward_stages import _conditioned_inputs
C_ = ref.load()
N, W, H, C = 40000, 480, 360, 1
means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 2024, W, H, C)
bg = torch.tensor([[1.0, 1.0, 1.0]], device="cuda")
for seed in range(8):
    torch.manual_seed(seed)
    leaves = [t.detach().clone().requires_grad_(True) for t in (means, covars, opac, betas, colors)]
    rc, ra, meta = ubs_b200.rasterization(leaves[0], None, None, leaves[2], leaves[3], leaves[4], viewmats, Ks, W, H, backgrounds=bg, covars=leaves[1])
    v_rc = torch.randn_like(rc) / (H * W); v_ra = torch.randn_like(ra) / (H * W)
    torch.autograd.backward((rc, ra), (v_rc, v_ra))
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc.contiguous(), v_ra.contiguous())
    tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    cov6 = covars[..., tri[0], tri[1]].contiguous()
    r_means, r_cov6, _, _, _ = C_.fully_fused_projection_bwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3, False, R["radii"], R["conics"], None, g2d, torch.zeros_like(R["depths"]), gcon, None, False)
    r_cov = torch.zeros(N, 3, 3, device="cuda"); r_cov[:, tri[0], tri[1]] = r_cov6
    # reference run-to-run noise: run its bwd twice
    g2d2, gcon2, *_ = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc.contiguous(), v_ra.contiguous())
    out = []
    for name, a, b in (("means", leaves[0].grad, r_means), ("covars", leaves[1].grad, r_cov), ("opac", leaves[2].grad, gop.sum(0)), ("betas", leaves[3].grad, gbe.sum(0)), ("colors", leaves[4].grad, gcol.sum(0))):
        scale = b.abs().max().clamp_min(1e-20)
        err = ((a - b).abs().max() / scale).item()
        big = b.abs() > 1e-3 * scale
        rel = ((a - b).abs() / b.abs().clamp_min(1e-30))[big]
        out.append("%s %.1e/%.4f%%" % (name, err, 100 * (rel > 3e-2).float().mean().item()))
    noise = ((gcon - gcon2).abs().max() / gcon.abs().max()).item()
    print(seed, " ".join(out), "ref-vs-ref conics noise %.1e" % noise)
