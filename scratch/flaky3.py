import sys, math, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import ref_cuda as ref
import ubs_b200
from test_gpu_forward_stages import _conditioned_inputs
C_ = ref.load()
N, W, H, C = 40000, 480, 360, 1
means, covars, opac, betas, colors, viewmats, Ks = _conditioned_inputs(N, 2024, W, H, C)
bg = torch.tensor([[1.0, 1.0, 1.0]], device="cuda")
tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
cov6 = covars[..., tri[0], tri[1]].contiguous()
torch.manual_seed(2)
leaves = [t.detach().clone().requires_grad_(True) for t in (means, covars, opac, betas, colors)]
rc, ra, meta = ubs_b200.rasterization(leaves[0], None, None, leaves[2], leaves[3], leaves[4], viewmats, Ks, W, H, backgrounds=bg, covars=leaves[1])
v_rc = torch.randn_like(rc) / (H * W); v_ra = torch.randn_like(ra) / (H * W)
torch.autograd.backward((rc, ra), (v_rc, v_ra))
R = ref.rasterization_fwd(means, covars, opac, betas, colors, viewmats, Ks, W, H, backgrounds=bg)
g2d, gcon, gcol, gop, gbe = C_.rasterize_to_pixels_bwd(R["means2d"], R["conics"], R["colors"], R["opacities"], R["betas"], bg, None, W, H, 16, R["isect_offsets"], R["flatten_ids"], R["render_alphas"], R["last_ids"], v_rc.contiguous(), v_ra.contiguous())
r_means = C_.fully_fused_projection_bwd(means, cov6, None, None, viewmats, Ks, W, H, 0.3, False, R["radii"], R["conics"], None, g2d, torch.zeros_like(R["depths"]), gcon, None, False)[0]
d = (leaves[0].grad - r_means).abs()
scale = r_means.abs().max()
i = int(d.max(dim=1).values.argmax())
print("worst prim", i, "err/scale %.2e" % (d.max() / scale).item(), "ours", leaves[0].grad[i].tolist(), "ref", r_means[i].tolist())
print(" means2d ours", meta["means2d"][0, i].tolist(), "ref", R["means2d"][0, i].tolist())
print(" conics ours", meta["conics"][0, i].tolist(), "ref", R["conics"][0, i].tolist())
print(" radii", meta["radii"][0, i].item(), R["radii"][0, i].item(), "depth", meta["depths"][0, i].item(), R["depths"][0, i].item())
print(" image diff max", (rc - R["render_colors"]).abs().max().item(), "last_ids differ", int((meta.get("last_ids", R["last_ids"]) != R["last_ids"]).sum()) if "last_ids" in meta else "n/a")
print(" covar", covars[i].tolist())
print(" g2d ref", g2d[0, i].tolist(), "gcon ref", gcon[0, i].tolist())
