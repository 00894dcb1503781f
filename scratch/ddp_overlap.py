"""2-GPU experiment: projection backward -> all-reduce -> Adam, sequential vs chunk-pipelined.
    torchrun --nproc-per-node 2 scratch/ddp_overlap.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth, training, parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
scene, cams, bg, cfg = synth.make_config("cfg3", device="cuda", cams_override=8)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
gt = torch.rand(1, 3, H, W, device="cuda")
if rank == 0: print("NCCL_DEBUG", os.environ.get("NCCL_DEBUG"), "HIGH_PRIO", os.environ.get("TORCH_NCCL_HIGH_PRIORITY"))
for chunks in (1, 2, 4, 8, 16):
    ts = training.TrainStep(rz, training.PackedAdam(scene.D, scene.N), world=world, n_chunks=chunks)
    r = rec.clone()
    def step(k):
        cam = cams[(k * world + rank) % len(cams)]
        ts.step(r, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None], gt, batch_size=world)
    for k in range(3): step(k)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20): step(k)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("chunks %2d: %.3f ms/step" % (chunks, e0.elapsed_time(e1) / 20))
dist.destroy_process_group()
