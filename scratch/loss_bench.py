import sys, torch
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import training
H, W = 1080, 1920
gt = torch.rand(1, 3, H, W, device="cuda")
img = torch.rand(1, H, W, 3, device="cuda")
v = torch.empty_like(img); out = torch.empty(3, device="cuda")
for _ in range(5): training.l1_ssim_loss_fwd_bwd(img, gt, 0.2, 1.0, "NHWC", "NCHW", True, v, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): training.l1_ssim_loss_fwd_bwd(img, gt, 0.2, 1.0, "NHWC", "NCHW", True, v, out)
e1.record(); torch.cuda.synchronize()
print("l1_ssim fwd+bwd 1080p: %.1f us" % (1e3 * e0.elapsed_time(e1) / 50))
