import sys, math, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import ref_cuda as ref
import ubs_b200
from ubs_b200 import ops
from test_gpu_forward_stages import _conditioned_inputs
C_ = ref.load()
N, W, H, C = 40000, 480, 360, 1
means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 2024, W, H, C)
bg = torch.tensor([[1.0, 1.0, 1.0]], device="cuda")
tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
cov6 = covars[..., tri[0], tri[1]].contiguous()
for seed in (2, 4):
    torch.manual_seed(seed)
    R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
    v_rc = torch.randn(1, H, W, 3, device="cuda") / (H * W); v_ra = torch.randn(1, H, W, 1, device="cuda") / (H * W)
    a = (R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc, v_ra)
    r = C_.rasterize_to_pixels_bwd(*a)
    m = ops.rasterize_bwd(*a)
    def proj(g2d, gcon):
        return C_.fully_fused_projection_bwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3, False, R["radii"], R["conics"], None, g2d.contiguous(), torch.zeros_like(R["depths"]), gcon.contiguous(), None, False)[0]
    pm_ref = proj(r[0], r[1]); pm_mine = proj(m[0], m[1])
    # our projection bwd on the reference's raster grads
    v_means_ours, v_cov_ours = ops.projection_bwd(means, cov6, viewmats, Ks, W, H, 0.3, R["radii"], R["conics"], None, r[0].contiguous(), torch.zeros_like(R["depths"]), r[1].contiguous(), None, False)[:2] if hasattr(ops, "projection_bwd") else (None, None)
    scale = pm_ref.abs().max()
    d = (pm_mine - pm_ref).abs()
    i = int(d.max(dim=1).values.argmax())
    print("seed", seed, "raster-stage induced err on means: %.2e of scale; worst prim %d" % ((d.max() / scale).item(), i))
    print("  ref v_means", pm_ref[i].tolist(), "mine-chain", pm_mine[i].tolist(), "scale", scale.item())
    print("  v_means2d ref", r[0][0, i].tolist(), "mine", m[0][0, i].tolist())
    print("  v_conics  ref", r[1][0, i].tolist(), "mine", m[1][0, i].tolist())
    print("  conic", R["conics"][0, i].tolist(), "radius", R["radii"][0, i].item(), "mean2d", R["means2d"][0, i].tolist(), "depth", R["depths"][0, i].item())
    if v_means_ours is not None:
        print("  proj-stage err (ours vs ref on same inputs): %.2e" % ((v_means_ours - pm_ref).abs().max() / scale).item())
    # reference fed with its own grads twice (noise floor through the chain)
    r2 = C_.rasterize_to_pixels_bwd(*a)
    print("  ref chain noise floor: %.2e" % ((proj(r2[0], r2[1]) - pm_ref).abs().max() / scale).item())
