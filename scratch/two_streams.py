"""Experiment: frames of different cameras on 1 / 2 / 3 streams, each with its own FusedRasterizer buffers."""
import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
scene, cams, bg, cfg = synth.make_config("cfg3", device="cuda", cams_override=16)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
for S in (1, 2, 3):
    rzs = [fused.FusedRasterizer(scene.D, scene.N, W, H, 1) for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    def frame(k):
        cam = cams[k % len(cams)]
        with torch.cuda.stream(streams[k % S]):
            rzs[k % S].forward(rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None])
    for k in range(6): frame(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    e0.record()
    for s in streams: s.wait_event(e0)
    for k in range(n): frame(k)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print("streams %d: %.3f ms/frame -> %.1f fps" % (S, e0.elapsed_time(e1) / n, 1000 * n / e0.elapsed_time(e1)))
