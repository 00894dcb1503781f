import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
scene, cams, bg, cfg = synth.make_config("cfg3", device="cuda", cams_override=4)
rec = fused.pack_records(scene.D, *scene.tensors())
rz = fused.FusedRasterizer(scene.D, scene.N, cfg["width"], cfg["height"], 1)
cam = cams[0]
rz.forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None])
n = rz.last_pair_count()
offs = torch.cat([rz.offsets.reshape(-1).long(), torch.tensor([n], device="cuda")])
cnt = offs[1:] - offs[:-1]
q = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0], device="cuda")
print("pairs/tile quantiles", torch.quantile(cnt.float(), q).tolist(), "mean", cnt.float().mean().item())
