"""Stage times of the full single-GPU train step (forward, loss, backward with the Adam epilogue)."""
import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth, training
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
fuse = not (len(sys.argv) > 2 and sys.argv[2] == "unfused")
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=16)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
tstep = training.TrainStep(rz, training.PackedAdam(scene.D, scene.N), fuse_adam=fuse)
gt = torch.rand(1, 3, H, W, device="cuda")
def step(k):
    cam = cams[k % len(cams)]
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    tstep.step(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None], gt, opacity_reg=0.01, scale_reg=0.01)
for k in range(4): step(k)
torch.cuda.synchronize()
rz.enable_stage_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(32): step(k)
e1.record(); torch.cuda.synchronize()
print(name, "%.3f ms/step" % (e0.elapsed_time(e1) / 32), {k: round(v[1], 4) for k, v in rz.stage_times_ms().items()})
