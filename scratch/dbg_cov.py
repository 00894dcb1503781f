import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from oracle import ref_cuda as ref
from ubs_b200 import ops, synth
C_ = ref.load()
for D in (6, 7):
    sc = synth.make_scene(20000, D, seed=5 + D).to("cuda")
    scale = torch.nn.functional.softplus(sc.scale)
    ri, rj = ref.tril_rest(D, "cuda")
    rot = C_.l_triangle_to_rotmat_fwd(sc.l_triangle[:, :3].contiguous())
    a = C_.rot_scale_l_triangle_to_covar_fwd(rot, scale, sc.l_triangle, ri, rj, False)
    b = ops.rot_scale_l_triangle_to_covar(rot, scale, sc.l_triangle, ri, rj, False)
    ne = (a != b)
    print(D, "mismatch count per (r,c):")
    print(ne.sum(0))
