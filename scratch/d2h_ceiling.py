"""Device->host copy ceiling of the box for the e2e path: every rank copies a 1920x1080x3 FP32 image (24.9 MB) from
device to pinned host memory in a loop, all ranks at once.  torchrun --nproc-per-node N scratch/d2h_ceiling.py [bind]"""
import os, sys, time, json
sys.path.insert(0, "universal-beta-splatting_b200")
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
info = {}
if bind:
    from ubs_b200 import hostmem
    info = hostmem.bind_to_gpu_node(lr)
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1920 * 1080 * 3
dev = [torch.rand(n, device="cuda") for _ in range(2)]
host = [torch.empty(n).pin_memory() for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
def run(iters, two_streams):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(iters):
        s = streams[k % 2] if two_streams else streams[0]
        with torch.cuda.stream(s):
            host[k % 2].copy_(dev[k % 2], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    return iters * n * 4 / t / 1e9
for two in (False, True):
    run(20, two)
    gbs = run(200, two)
    t = torch.tensor([gbs], device="cuda")
    tot = t.clone()
    if world > 1:
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    else:
        lo = t
    if rank == 0:
        print(json.dumps({"ranks": world, "bind": bind, "two_streams": two, "GBs_per_rank_min": round(lo.item(), 2),
                          "GBs_aggregate": round(tot.item(), 2), "frames_per_s_ceiling": round(tot.item() * 1e9 / (n * 4), 1),
                          "rank0_info": info}))
if world > 1: dist.destroy_process_group()
