"""Generates tests/golden/loss_ssim.npz from the REFERENCE's own utils/loss_utils.py (pure torch, imported from
/root/reference in the build container; it does not exist on the GPU box, hence the committed fixture).

    python tests/golden/make_golden_loss.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for tag, (ch, H, W), seed in (("a", (3, 37, 53), 11), ("b", (3, 64, 96), 12), ("c", (1, 19, 8), 13)):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand((ch, H, W), generator=g)
    img = (gt + 0.15 * torch.randn((ch, H, W), generator=g)).clamp(0, 1)
    if tag == "b":
        img[:, :20, :30] = gt[:, :20, :30]  # exact-match region: sign(0) = 0 in the L1 gradient
    img.requires_grad_(True)
    l1 = ref.l1_loss(img, gt)
    s = ref.ssim(img.unsqueeze(0), gt.unsqueeze(0))
    lam = 0.2
    loss = (1.0 - lam) * l1 + lam * (1.0 - s)  # train.py:118-121 with loss_utils.ssim in place of fused_ssim
    loss.backward()
    out.update({f"{tag}_img": img.detach().numpy(), f"{tag}_gt": gt.numpy(), f"{tag}_l1": l1.item(),
                f"{tag}_ssim": s.item(), f"{tag}_loss": loss.item(), f"{tag}_grad": img.grad.numpy()})
np.savez_compressed(os.path.join(HERE, "loss_ssim.npz"), **out)
print({k: v for k, v in out.items() if np.ndim(v) == 0})
