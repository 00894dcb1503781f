"""A/B of the compositing-forward variants: python scratch/ab_fwd.py <cfg> ; set UBS_FWD_VARIANT / UBS_BWD_VARIANT.
Prints stage times and checksums; saves / compares the image of camera 0 across variants via /tmp/ab_fwd_<cfg>.pt."""
import os, sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=8)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
def args(c):
    ts = torch.tensor([c.timestamp], device="cuda") if scene.D == 7 else None
    return (rec, c.viewmat[None], c.K[None], c.cam_pos[None], ts, bg[None])
P = W * H
g = torch.Generator(device="cuda").manual_seed(1)
v_rc = torch.randn(1, H, W, 3, device="cuda", generator=g) / P
v_ra = torch.zeros(1, H, W, 1, device="cuda")
vrec = torch.empty_like(rec)
def step(k):
    a = args(cams[k % len(cams)])
    rz.forward(*a)
    rz.backward(*a, v_rc, v_ra, vrec)
for k in range(4): step(k)
torch.cuda.synchronize()
rz.enable_stage_timing(True)
for k in range(24): step(k)
st = rz.stage_times_ms()
tag = "fwd_variant=%s bwd_variant=%s" % (os.environ.get("UBS_FWD_VARIANT", "0"), os.environ.get("UBS_BWD_VARIANT", "0"))
print("%s %s: %s" % (name, tag, {k: round(v[1], 4) for k, v in st.items()}))
rz.enable_stage_timing(False)
a = args(cams[0])
rc, ra = rz.forward(*a)
rz.backward(*a, v_rc, v_ra, vrec)
torch.cuda.synchronize()
path = "/tmp/ab_fwd_%s.pt" % name
cur = dict(rc=rc.clone().cpu(), ra=ra.clone().cpu(), last=rz.last_ids.clone().cpu(), g=vrec.clone().cpu())
if os.path.exists(path):
    ref = torch.load(path)
    gs = ref["g"].abs().max(dim=0).values.clamp_min(1e-30)
    print("  vs first variant: image equal %s, alpha equal %s, last_ids equal %s, grad max rel-to-column-scale diff %.3e"
          % (torch.equal(ref["rc"], cur["rc"]), torch.equal(ref["ra"], cur["ra"]), torch.equal(ref["last"], cur["last"]),
             float(((ref["g"] - cur["g"]).abs() / gs).max())))
else:
    torch.save(cur, path)
