import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
cam = cams[0]
rec = fused.pack_records(scene.D, *scene.tensors())
rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1)
ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
P = cam.width * cam.height
v_rc = torch.randn(1, cam.height, cam.width, 3, device="cuda") / P
v_ra = torch.zeros(1, cam.height, cam.width, 1, device="cuda")
vrec = torch.empty_like(rec)
def step():
    rz.forward(*args)
    rz.backward(*args, v_rc, v_ra, vrec)
for _ in range(3): step()
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
rz.enable_stage_timing(True)
for _ in range(n): step()
st = rz.stage_times_ms()
print("%s: fwd+bwd %.3f ms/iter -> %.1f it/s; stages %s" % (name, e0.elapsed_time(e1) / n, 1000 * n / e0.elapsed_time(e1), {k: round(v[1], 4) for k, v in st.items()}))
print("grad checksum", float(vrec.double().abs().sum()))
