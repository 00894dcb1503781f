"""Where do the 48-byte gradient rows (ubs_rasterize_bwd_rows) differ from the separate arrays (ubs_rasterize_bwd_splats)?
Prints, per screen-space quantity, the worst primitives with their tile counts, for one camera of a config."""
import math, os, sys
import torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
cam = cams[0]
rec = fused.pack_records(scene.D, *scene.tensors())
ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
P = cam.width * cam.height
g = torch.Generator(device="cuda").manual_seed(3)
v_rc = torch.randn(1, cam.height, cam.width, 3, device="cuda", generator=g) / P
v_ra = torch.randn(1, cam.height, cam.width, 1, device="cuda", generator=g) / P
res = {}
for key, rows in (("rows", True), ("arrays", False), ("arrays2", False)):
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1, grad_rows=rows)
    rz.forward(*args)
    vrec = rz.backward(*args, v_rc, v_ra)
    if rows:
        r = rz.v_rows[0].double()
        a, b, c = rz.conics[0].double().unbind(-1)
        d = {"v_colors": r[:, 0:3], "v_conics": torch.stack((r[:, 3], 2 * r[:, 4], r[:, 5]), -1),
             "v_means2d": torch.stack((2 * a * r[:, 6] + 2 * b * r[:, 7], 2 * b * r[:, 6] + 2 * c * r[:, 7]), -1),
             "v_opacities": r[:, 8:9], "v_betas": r[:, 9:10] * math.log(2.0)}
    else:
        d = {"v_colors": rz.v_colors[0].double(), "v_conics": rz.v_conics[0].double(), "v_means2d": rz.v_means2d[0].double(),
             "v_opacities": rz.v_opacities[0].double()[:, None], "v_betas": rz.v_betas[0].double()[:, None]}
    d["rec"] = vrec.double()
    res[key] = d
    tiles = rz.tiles_per_gauss[0].clone(); radii = rz.radii[0].clone()
    del rz
for q in res["rows"]:
    x, y, y2 = res["rows"][q], res["arrays"][q], res["arrays2"][q]
    scale = y.abs().max().item()
    e = (x - y).abs().amax(dim=1); e2 = (y2 - y).abs().amax(dim=1)
    print("%-12s scale %.3e  rows-arrays max %.3e  arrays-arrays max %.3e" % (q, scale, e.max().item(), e2.max().item()))
    for i in e.topk(3).indices.tolist():
        print("    prim %8d tiles %6d radius %5d  rows %s  arrays %s  arrays2 %s" % (
            i, tiles[i].item(), radii[i].item(), x[i].tolist()[:3], y[i].tolist()[:3], y2[i].tolist()[:3]))
a, b, b2 = res["rows"]["rec"], res["arrays"]["rec"], res["arrays2"]["rec"]
scale = b.abs().amax(dim=0).clamp_min(1e-30)
print("per column rows-arrays  ", ["%.1e" % v for v in ((a - b).abs().amax(dim=0) / scale).tolist()])
print("per column arrays-arrays", ["%.1e" % v for v in ((b2 - b).abs().amax(dim=0) / scale).tolist()])
col = int(((a - b).abs().amax(dim=0) / scale).argmax())
i = int((a - b).abs()[:, col].argmax())
print("worst column", col, "prim", i, "tiles", tiles[i].item(), "radius", radii[i].item(), a[i, col].item(), b[i, col].item(), b2[i, col].item())
for q in ("v_colors", "v_conics", "v_means2d", "v_opacities", "v_betas"):
    print("   ", q, res["rows"][q][i].tolist(), res["arrays"][q][i].tolist())
