"""Short driver for ncu captures: a few forward+backward frames of the bench workload (cfg3, 3M prims, 1080p).

    ncu --set full --clock-control none --import-source on -k regex:rasterize_fwd -s 3 -c 1 \
        -o gpurun_out/prof_rasterize_fwd python profiles/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-beta-splatting_b200"))
from ubs_b200 import fused, synth, training  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
full = len(sys.argv) > 3 and sys.argv[3].startswith("full")  # full train step: + L1/SSIM loss + Adam
fuse = not (len(sys.argv) > 3 and sys.argv[3] == "full_unfused")  # Adam inside the projection backward (1 GPU) or separate
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=8)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
v_rc = torch.randn(1, H, W, 3, device="cuda") / (W * H)
v_ra = torch.zeros(1, H, W, 1, device="cuda")
vrec = torch.empty_like(rec)
bgd = bg[None]
if full:
    tstep = training.TrainStep(rz, training.PackedAdam(scene.D, scene.N), fuse_adam=fuse)
    gt = torch.rand(1, 3, H, W, device="cuda")
for k in range(iters):
    cam = cams[k % len(cams)]
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bgd)
    if full:
        tstep.step(*args, gt, opacity_reg=0.01, scale_reg=0.01)
        continue
    rz.forward(*args)
    rz.backward(*args, v_rc, v_ra, vrec)
torch.cuda.synchronize()
print("pairs", rz.last_pair_count())
