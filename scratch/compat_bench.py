"""Drop-in path: the reference's own call sequence (scene/beta_model.py:660-711) through the shim's operators
(ubs_b200.ops + rasterization()) with torch autograd, against the fused fast path.  cfg3."""
import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
import ubs_b200
from ubs_b200 import ops, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=8)
D, N, W, H = scene.D, scene.N, cfg["width"], cfg["height"]
params = [t.clone().requires_grad_(True) for t in scene.tensors()]
ti, tj = torch.tril_indices(D, D, offset=-1)
m = (ti >= 3) | (tj >= 3)
rest_i, rest_j = ti[m].int().cuda(), tj[m].int().cuda()
def render(cam):
    xyz, mean, rgb, opacity, beta, scale, ltri = params
    s = torch.nn.functional.softplus(scale); o = torch.sigmoid(opacity); b = 4.0 * torch.exp(beta)
    rot = ops.l_triangle_to_rotmat(ltri[:, :3].contiguous())
    cov = ops.rot_scale_l_triangle_to_covar(rot, s, ltri, rest_i, rest_j)
    vd = xyz - cam.cam_pos[None]; vd = vd / vd.norm(dim=-1, keepdim=True)
    q = vd if D == 6 else torch.cat([vd, torch.full((N, 1), cam.timestamp, device="cuda")], -1)
    means, covs, opac = ops.cond_mean_convariance_opacity(torch.cat([xyz, mean], -1), cov, o, b[:, 1:].contiguous(), q.detach())
    return ubs_b200.rasterization(means, ltri, s, opac.squeeze(-1), b[:, 0], rgb, cam.viewmat[None], cam.K[None], W, H,
                                  backgrounds=bg[None], covars=covs)
v = torch.randn(1, H, W, 3, device="cuda") / (W * H)
def fwd(k):
    with torch.no_grad(): render(cams[k % 8])
def fwdbwd(k):
    rc, ra, meta = render(cams[k % 8])
    for p in params: p.grad = None
    rc.backward(v)
for nm, fn in (("forward", fwd), ("forward+backward", fwdbwd)):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10): fn(k)
    e1.record(); torch.cuda.synchronize()
    print("%s drop-in path (shim ops + rasterization()), %s: %.3f ms" % (name, nm, e0.elapsed_time(e1) / 10))
