"""Per-stage CUDA-event times of forward (+ backward) over several cameras of the bench workload.
    python scratch/stage_bench.py cfg3 [bwd]"""
import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
bwd = len(sys.argv) > 2 and sys.argv[2] == "bwd"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=16)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
v_rc = torch.randn(1, H, W, 3, device="cuda") / (W * H)
v_ra = torch.zeros(1, H, W, 1, device="cuda")
vrec = torch.empty_like(rec)
def step(k):
    cam = cams[k % len(cams)]
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    rz.forward(*args)
    if bwd: rz.backward(*args, v_rc, v_ra, vrec)
for k in range(4): step(k)
rz.enable_stage_timing(True)
for k in range(32): step(k)
print(name, {k: round(v[1], 4) for k, v in rz.stage_times_ms().items()})
