"""Repeatability of the gradient-rows backward: K backward passes of the same frame, each compared with the separate-array result."""
import sys
import torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
for name in sys.argv[1:] or ["cfg4"]:
    scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
    cam = cams[0]
    rec = fused.pack_records(scene.D, *scene.tensors())
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    P = cam.width * cam.height
    g = torch.Generator(device="cuda").manual_seed(3)
    v_rc = torch.randn(1, cam.height, cam.width, 3, device="cuda", generator=g) / P
    v_ra = torch.randn(1, cam.height, cam.width, 1, device="cuda", generator=g) / P
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1, grad_rows=False)
    rz.forward(*args); base = rz.backward(*args, v_rc, v_ra).double()
    scale = base.abs().amax(dim=0).clamp_min(1e-30)
    del rz
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1, grad_rows=True)
    errs = []
    for it in range(30):
        rz.forward(*args)
        v = rz.backward(*args, v_rc, v_ra).double()
        errs.append(float(((v - base).abs().amax(dim=0) / scale).max()))
    print(name, "pairs", rz.last_pair_count(), "max column error over 30 passes:", ["%.1e" % e for e in errs], flush=True)
    del rz
