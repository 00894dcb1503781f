"""Small forward + backward (rows and arrays) for compute-sanitizer (initcheck / racecheck / memcheck)."""
import sys
import torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
D, N, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 7, 20000, 333, 250
scene = synth.make_scene(N, D, seed=5).to("cuda")
cam = synth.make_cameras(1, W, H, seed=6, timestamps=[0.3], device="cuda")[0]
rec = fused.pack_records(D, *scene.tensors())
ts = torch.tensor([cam.timestamp], device="cuda") if D == 7 else None
bg = torch.rand(1, 3, device="cuda")
args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg)
v_rc = torch.randn(1, H, W, 3, device="cuda") / (H * W)
v_ra = torch.randn(1, H, W, 1, device="cuda") / (H * W)
for rows in (True, False):
    rz = fused.FusedRasterizer(D, N, W, H, 1, grad_rows=rows)
    for _ in range(2):
        rz.forward(*args)
        v = rz.backward(*args, v_rc, v_ra)
    torch.cuda.synchronize()
    print("rows" if rows else "arrays", float(v.abs().sum()), rz.last_pair_count())
