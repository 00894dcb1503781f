import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
for name in sys.argv[1:] or ["cfg3"]:
    scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
    cam = cams[0]
    rec = fused.pack_records(scene.D, *scene.tensors())
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1)
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    c = rz.work_counts()
    print(name, {k: ("%.4g" % v) for k, v in c.items()})
    off = rz.offsets.flatten().long()
    n = int(rz.n_isects.item())
    seg = torch.diff(torch.cat([off, torch.tensor([n], device="cuda")]))
    print(name, "pairs/tile: mean %.1f  median %d  p99 %d  max %d ; radii mean %.2f median %d p99 %d max %d" % (
        seg.float().mean().item(), seg.median().item(), seg.float().quantile(0.99).item(), seg.max().item(),
        rz.radii[rz.radii > 0].float().mean().item(), rz.radii[rz.radii > 0].median().item(),
        rz.radii[rz.radii > 0].float().quantile(0.99).item(), rz.radii.max().item()))
