"""A/B of the RGB compositing-backward kernels (variant 0: lane per pixel + butterfly, 1: lane per pair + scan, both into
the separate gradient arrays; "rows": lane per pair, vector reductions into 48-byte gradient rows = the fused default).
Runs each variant in its own process, compares the packed gradient records and prints the stage times."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
    from ubs_b200 import fused, synth
    name = sys.argv[2]
    scene, cams, bg, cfg = synth.make_config(name, device="cuda", cams_override=1)
    cam = cams[0]
    rec = fused.pack_records(scene.D, *scene.tensors())
    rz = fused.FusedRasterizer(scene.D, scene.N, cam.width, cam.height, 1, grad_rows=os.environ.get("AB_ROWS") == "1")
    ts = torch.tensor([cam.timestamp], device="cuda") if scene.D == 7 else None
    args = (rec, cam.viewmat[None], cam.K[None], cam.cam_pos[None], ts, bg[None])
    P = cam.width * cam.height
    g = torch.Generator(device="cuda").manual_seed(3)
    v_rc = torch.randn(1, cam.height, cam.width, 3, device="cuda", generator=g) / P
    v_ra = torch.randn(1, cam.height, cam.width, 1, device="cuda", generator=g) / P
    vrec = torch.empty_like(rec)
    def step():
        rz.forward(*args)
        rz.backward(*args, v_rc, v_ra, vrec)
    for _ in range(3): step()
    torch.cuda.synchronize()
    rz.enable_stage_timing(True)
    for _ in range(10): step()
    st = rz.stage_times_ms()
    print(name, "variant", os.environ.get("AB_NAME"), {k: round(v[1], 4) for k, v in st.items()}, flush=True)
    torch.save(vrec.cpu(), sys.argv[3])
else:
    import torch
    for name in sys.argv[1:] or ["cfg3"]:
        outs = []
        for v in os.environ.get("AB_VARIANTS", "0,1,rows").split(","):
            f = "/tmp/ab_bwd3_%s_%s.pt" % (name, v)
            subprocess.run([sys.executable, __file__, "child", name, f], env=dict(os.environ, UBS_BWD3_VARIANT={"rows": "1"}.get(v, v), AB_ROWS="1" if v == "rows" else "0", AB_NAME=v), check=True)
            outs.append(torch.load(f).double())
        a, b = outs[0], outs[-1]
        scale = a.abs().amax(dim=0).clamp_min(1e-30)
        err = ((a - b).abs().amax(dim=0) / scale)
        print(name, "max |v0 - v1| / max |v0| per record column:", [float("%.2e" % x) for x in err.tolist()])
