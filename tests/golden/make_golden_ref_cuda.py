"""Generates tests/golden/ref_cuda_D{6,7}.npz by RUNNING the reference's own compiled CUDA kernels
(oracle/_ref/ubs_ref_cuda.so, built from /root/reference by oracle/build_ref.py) on a B200:

    gpurun -- 'python tests/golden/make_golden_ref_cuda.py'     # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/ref_cuda_D*.npz tests/golden/

The fixtures hold, for one small seeded scene per D: every intermediate of the forward chain
K1->K2->K3->K5->K7..K9->K10 (scene/beta_model.py:660-711 call order) and of the backward chain
K11->K6->K4->K2bwd->K1bwd, down to the gradients of the 7 raw parameter tensors.  They pin the CPU oracle
(tests/test_oracle_golden.py, no GPU needed) and the CUDA library (tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "universal-beta-splatting_b200"))
sys.path.insert(0, ROOT)

from oracle import ref_cuda as ref  # noqa: E402
from ubs_b200 import synth  # noqa: E402

SCENES = {6: dict(N=3000, W=176, H=112, seed=606, ts=0.0), 7: dict(N=2500, W=144, H=96, seed=707, ts=0.4)}


def scene_and_camera(D, device="cpu"):
    """Shared by the generator and by the tests that replay the fixture (inputs are regenerated from the seed)."""
    s = SCENES[D]
    scene = synth.make_scene(s["N"], D, seed=s["seed"], extent=5.0)
    # enlarge the footprints (x3) so that a few thousand primitives give deep per-pixel lists and early termination
    scene.scale[:, :3] = synth.inverse_softplus(torch.nn.functional.softplus(scene.scale[:, :3]) * 3.0)
    cam = synth.make_cameras(1, s["W"], s["H"], radius=6.0, seed=s["seed"], timestamps=[s["ts"]])[0]
    bg = torch.tensor([0.2, 0.5, 0.9])
    g = torch.Generator().manual_seed(s["seed"] + 1)
    P = s["W"] * s["H"]
    v_rc = torch.randn(1, s["H"], s["W"], 3, generator=g) / P
    v_ra = torch.randn(1, s["H"], s["W"], 1, generator=g) / P
    mv = lambda t: t.to(device)  # noqa: E731
    scene = scene.to(device)
    cam = synth.Camera(mv(cam.viewmat), mv(cam.K), mv(cam.cam_pos), cam.width, cam.height, cam.timestamp)
    return scene, cam, mv(bg), mv(v_rc), mv(v_ra)


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for D in (6, 7):
        scene, cam, bg, v_rc, v_ra = scene_and_camera(D, "cuda")
        keep = {}
        grads, R = ref.chain_grads(scene, cam, bg, v_rc, v_ra, keep=keep)
        out = {("fwd_" + k): v for k, v in R.items() if v is not None}
        out.update({("mid_" + k): v for k, v in keep.items()})
        for name, g in zip(("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle"), grads):
            out["grad_" + name] = g
        npz = {k: v.detach().cpu().numpy() for k, v in out.items()}
        vis = int((R["radii"] > 0).sum())
        print("D=%d visible %d/%d pairs %d alpha>0.5 %.3f max last_id %d" % (
            D, vis, scene.N, R["isect_ids"].numel(), (R["render_alphas"] > 0.5).float().mean().item(),
            int(R["last_ids"].max())))
        np.savez_compressed(os.path.join(out_dir, "ref_cuda_D%d.npz" % D), **npz)
        print("wrote", os.path.join(out_dir, "ref_cuda_D%d.npz" % D), os.path.getsize(
            os.path.join(out_dir, "ref_cuda_D%d.npz" % D)), "bytes")


if __name__ == "__main__":
    main()
