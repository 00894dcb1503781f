"""Where the zero-edit drop-in's time goes: torch profiler over BetaModel.render's statements (tests/ref_caller.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-beta-splatting_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from ubs_b200 import synth
import ref_caller
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=4)
model = ref_caller.BetaModelCaller(scene, bg, requires_grad=True)
vcs = [ref_caller.ViewpointCamera(c) for c in cams]
W, H = cfg["width"], cfg["height"]
v_img = torch.randn(3, H, W, device="cuda") / (W * H)
def fwd(k):
    with torch.no_grad():
        model.render(vcs[k % 4])
def train(k):
    out = model.render(vcs[k % 4])
    (out["render"] * v_img).sum().backward()
    for t in model.leaves(): t.grad = None
for fn, nm in ((fwd, "fwd"), (train, "fwd+bwd")):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10): fn(k)
    e1.record(); torch.cuda.synchronize()
    print("==== %s: %.3f ms/iter" % (nm, e0.elapsed_time(e1) / 10))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for k in range(5): fn(k)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
