"""N-GPU experiment: sharded P2P train step, parameter broadcast by peer stores vs NVLS multicast stores."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, "universal-beta-splatting_b200"); sys.path.insert(0, ".")
from ubs_b200 import fused, synth, training, parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
parallel.init_process_group("nccl", device_id=torch.device("cuda", rank))
scene, cams, bg, cfg = synth.make_config("cfg3", device="cuda", cams_override=64)
rec = fused.pack_records(scene.D, *scene.tensors())
W, H = cfg["width"], cfg["height"]
rz = fused.FusedRasterizer(scene.D, scene.N, W, H, 1)
gt = torch.rand(1, 3, H, W, device="cuda")
st = parallel.ShardedState.create(scene.D, scene.N)
st.records.copy_(rec)
mc = st.mc_records
if rank == 0: print("multicast ptr", mc)
for label, use in (("peer stores", None), ("multicast", mc), ("peer stores", None), ("multicast", mc)):
    st.mc_records = use
    ts = training.TrainStep(rz, training.PackedAdam(scene.D, scene.N, allocate_moments=False), world=world, sharded=st)
    def step(k):
        cam = cams[(k * world + rank) % len(cams)]
        ts.step(st.records, cam.viewmat[None], cam.K[None], cam.cam_pos[None], None, bg[None], gt, batch_size=world)
    for k in range(3): step(k)
    dist.barrier(); torch.cuda.synchronize()
    rz.enable_stage_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20): step(k)
    e1.record(); torch.cuda.synchronize()
    stt = rz.stage_times_ms(); rz.enable_stage_timing(False)
    if rank == 0: print("%-12s %.3f ms/step  scatter %.3f  reduce_adam_gather %.3f  barrier(avg of 2) %.3f" % (
        label, e0.elapsed_time(e1) / 20, stt["bwd_scatter"][1], stt["reduce_adam_gather"][1], stt["barrier"][1]))
dist.destroy_process_group()
