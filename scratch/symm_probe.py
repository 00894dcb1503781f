import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm_mem.empty((1024, 36), dtype=torch.float32, device=torch.device("cuda", rank))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "mc_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None)
t.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1024, 36), torch.float32)
peer[rank].fill_(100.0 + rank)      # P2P store into the peer's buffer
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "row values", t[:3, 0].tolist())
big = symm_mem.empty((3_000_064 * 36,), dtype=torch.float32, device=torch.device("cuda", rank))
h2 = symm_mem.rendezvous(big, dist.group.WORLD)
print(rank, "big ok", big.numel() * 4 / 1e6, "MB")
dist.destroy_process_group()
