import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bin", "onesweep"]
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
cam = cams[0]
rec = fused.pack_records(scene.D, *scene.tensors())
ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
for mode in modes:
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1, sort_mode=mode)
    args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    for _ in range(3): rz.forward(*args)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): rz.forward(*args)
    e1.record(); torch.cuda.synchronize()
    rz.enable_stage_timing(True)
    for _ in range(n): rz.forward(*args)
    st = rz.stage_times_ms()
    print("%s %s: %.3f ms/frame; stages %s" % (name, mode, e0.elapsed_time(e1) / n, {k: round(v[1], 4) for k, v in st.items()}))
    del rz
