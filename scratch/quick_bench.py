import sys, time, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
cam = cams[0]
rec = fused.pack_records(scene.D, *scene.tensors())
rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1)
ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
for _ in range(3): rz.forward(*args)
torch.cuda.synchronize()
print("pairs", rz.last_pair_count(), "visible", int((rz.radii > 0).sum()), "cap", rz.capacity, "overflow", rz.overflowed())
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): rz.forward(*args)
e1.record(); torch.cuda.synchronize()
print("%s: %.3f ms/frame -> %.1f fps" % (name, e0.elapsed_time(e1) / n, 1000 * n / e0.elapsed_time(e1)))
print("alpha>0.5 frac", (rz.render_alphas > 0.5).float().mean().item())
