#!/usr/bin/env python
"""Benchmark of the hot path: rendered frames/s (and forward+backward it/s) at 3M 6-D Beta primitives, 1920x1080.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this library (CUDA, sm_100a)
    python bench.py --impl reference ...                                # the CPU oracle port on the host cores
    python bench.py --impl reference_cuda ...                           # the reference's own CUDA kernels (if built)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU (camera-parallel)

One "step" = one camera rendered per GPU (a different camera of a 64-camera ring every step and rank); the scene
(parameter records, 432 MB) is resident in HBM and is larger than L2, so no L2 flush is needed between steps.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "universal-beta-splatting_b200"))
sys.path.insert(0, ROOT)

if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    # at these two levels NCCL prints its version banner to stdout, which carries the ONE JSON line; INFO and above
    # (e.g. when someone wants the NVLS evidence) are left alone
    os.environ["NCCL_DEBUG"] = "NONE"
# the gradient all-reduce is pipelined against the projection backward / Adam kernels (parallel.pipelined_backward);
# NCCL's kernels only get SMs next to those full-machine grids from a high-priority stream (measured at 2 GPUs)
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")

import torch  # noqa: E402

WORKLOADS = {
    # name -> (synth config, description)
    "cfg3": ("cfg3", "BASELINE configs[2]/[4]: 3M 6-D Beta primitives (unbounded), 1920x1080, 64-camera ring, "
                     "one camera per GPU per step"),
    "cfg3_r4": ("cfg3_r4", "BASELINE configs[2] at -r 4: 3M 6-D primitives, 1245x825"),
    "cfg2": ("cfg2", "BASELINE configs[1]: 300k 6-D primitives, 800x800, white background"),
    "cfg1": ("cfg1", "BASELINE configs[0]: 100k 6-D primitives, 800x800"),
    "cfg4": ("cfg4", "BASELINE configs[3]: 1M 7-D primitives, 1352x1014, 300 timestamps"),
}
N_RING = 64


# ----------------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.timeline = []  # (perf_counter, sm_mhz, reason bits) for region() queries
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # pragma: no cover
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                self.samples.append(mhz)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.timeline.append((time.perf_counter(), mhz, r))
                self.names = names
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # pragma: no cover
                pass
            time.sleep(self.period)

    def region(self, t0, t1):
        """Clocks line of the samples taken in [t0, t1] (perf_counter)."""
        sel = [x for x in list(self.timeline) if t0 <= x[0] <= t1]
        s = sorted(x[1] for x in sel)
        reasons = set()
        for _, _, r in sel:
            for bit, name in getattr(self, "names", {}).items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(s)}

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name, n_override=None):
    from ubs_b200 import synth

    cfg_name, desc = WORKLOADS[name]
    scene, cams, bg, cfg = synth.make_config(cfg_name, n_override=n_override, cams_override=N_RING)
    return scene, cams, bg, cfg, desc


# ----------------------------------------------------------------------------------------------------------------
# the reference arm: CPU oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_pass(scene, cam, bg, backward):
    """One frame through the CPU restatement of the reference path (oracle/): returns seconds."""
    from oracle import ubs_oracle as O

    t0 = time.perf_counter()
    if backward:
        params = [t.clone().requires_grad_(True) for t in scene.tensors()]
        rc, ra, _ = O.render(params, cam.viewmat, cam.K, cam.cam_pos, cam.timestamp, cam.width, cam.height, bg)
        P = cam.width * cam.height
        g = torch.Generator().manual_seed(1)
        torch.autograd.backward((rc, ra), (torch.randn(rc.shape, generator=g) / P, torch.zeros_like(ra)))
    else:
        with torch.no_grad():
            O.render(scene.tensors(), cam.viewmat, cam.K, cam.cam_pos, cam.timestamp, cam.width, cam.height, bg)
    return time.perf_counter() - t0


def run_reference_cpu(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    scene, cams, bg, cfg, desc = make_workload(args.workload)
    from oracle import ubs_oracle as O

    O.build()
    for w in range(min(args.warmup, 1)):
        cpu_oracle_pass(scene, cams[w % N_RING], bg, False)
    steps = max(1, min(args.steps, 5))  # bounded sample: each step is one whole frame, ~4 s on 8 cores
    times = [cpu_oracle_pass(scene, cams[(k + 1) % N_RING], bg, False) for k in range(steps)]
    t_bwd = cpu_oracle_pass(scene, cams[0], bg, True)
    fps = steps / sum(times)
    cores = os.cpu_count() or 1
    sample = "%d whole frames of the same workload (forward), 1 forward+backward frame" % steps
    line = {
        "impl": "reference", "metric": "render_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": 0,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "N": scene.N, "D": scene.D, "width": cfg["width"],
                   "height": cfg["height"]},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "train_it_per_s": 1.0 / t_bwd},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "train": {"value": 1.0 / t_bwd, "unit": "it/s", "ms_per_step": 1e3 * t_bwd},
        "note": "the reference's path is CUDA-only past projection (its torch rasteriser needs CUDA + nerfacc); "
                "this arm is the CPU restatement in oracle/ (torch per-primitive stages + C/OpenMP tile/compositing)",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# extra arm: the reference's own CUDA kernels on the same GPU (the number to beat)
# ----------------------------------------------------------------------------------------------------------------
def run_reference_cuda(args, rank, world):
    if rank != 0:
        return
    from oracle import ref_cuda as ref

    if not ref.available():
        print(json.dumps({"impl": "reference_cuda", "unavailable": "oracle/_ref/ubs_ref_cuda.so not built"}))
        return
    dev = "cuda:0"
    scene, cams, bg, cfg, desc = make_workload(args.workload)
    scene = scene.to(dev)
    bg = bg.to(dev)
    from ubs_b200 import synth

    cams = [synth.Camera(c.viewmat.to(dev), c.K.to(dev), c.cam_pos.to(dev), c.width, c.height, c.timestamp)
            for c in cams]
    W, H = cfg["width"], cfg["height"]
    P = W * H

    def fwd(k):
        cam = cams[k % N_RING]
        m, v, o, b0 = ref.condition(scene, cam)
        return ref.rasterization_fwd(m, v, o, b0, scene.rgb, cam.viewmat[None], cam.K[None], W, H,
                                     backgrounds=bg[None])

    v_rc = torch.randn(1, H, W, 3, device=dev) / P
    v_ra = torch.zeros(1, H, W, 1, device=dev)

    def train(k):
        ref.chain_grads(scene, cams[k % N_RING], bg, v_rc, v_ra)

    out = {}
    for name, fn in (("render", fwd), ("train", train)):
        for w in range(args.warmup):
            fn(w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            fn(args.warmup + k)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / args.steps
    line = {"impl": "reference_cuda", "metric": "render_frames_per_s", "value": 1e3 / out["render"],
            "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": out["render"], "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": args.workload},
            "train": {"value": 1e3 / out["train"], "unit": "it/s", "ms_per_step": out["train"]},
            "note": "reference gsplat CUDA kernels recompiled for sm_100a (oracle/_ref), driven as "
                    "scene/beta_model.py:660-711 drives them, on the same GPU"}
    print(json.dumps(line))


def _time_ms(fn, steps, warmup):
    for w in range(warmup):
        fn(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn(warmup + k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def measure_reference_cuda_and_dropin(scene, cams, bg, cfg, dev, our_its, our_fps, steps=10, warmup=3):
    """Rank 0, one GPU, outside every timed region of the headline numbers:
    reference_cuda -- the reference's own kernels (oracle/_ref, recompiled for sm_100a) driven as
                      scene/beta_model.py:660-711 drives them: frames/s, forward+backward it/s and our ratios to them;
    dropin         -- the reference CALLER's statements (tests/ref_caller.py, AST-identical to BetaModel.render)
                      against the gsplat shim: what a user who only swaps the import gets."""
    from ubs_b200 import dropin, synth

    W, H = cfg["width"], cfg["height"]
    scene_d = scene.to(dev)
    bg_d = bg.to(dev)
    cams_d = [synth.Camera(c.viewmat.to(dev), c.K.to(dev), c.cam_pos.to(dev), c.width, c.height, c.timestamp)
              for c in cams[:16]]
    g = torch.Generator(device=dev).manual_seed(3)
    v_img = torch.randn(3, H, W, device=dev, generator=g) / (W * H)
    ref_line = {"unavailable": "oracle/_ref/ubs_ref_cuda.so not built"}
    try:
        from oracle import ref_cuda as ref

        if ref.available():
            v_rc = v_img.permute(1, 2, 0)[None].contiguous()
            v_ra = torch.zeros(1, H, W, 1, device=dev)

            def r_fwd(k):
                cam = cams_d[k % len(cams_d)]
                m, v, o, b0 = ref.condition(scene_d, cam)
                ref.rasterization_fwd(m, v, o, b0, scene_d.rgb, cam.viewmat[None], cam.K[None], W, H,
                                      backgrounds=bg_d[None])

            ms_f = _time_ms(r_fwd, steps, warmup)
            ms_t = _time_ms(lambda k: ref.chain_grads(scene_d, cams_d[k % len(cams_d)], bg_d, v_rc, v_ra), steps, warmup)
            ref_line = {"frames_per_s": 1e3 / ms_f, "train_it_per_s": 1e3 / ms_t, "ms_per_frame": ms_f,
                        "ms_per_fwdbwd": ms_t, "ratio": our_fps / (1e3 / ms_f), "train_ratio": our_its / (1e3 / ms_t),
                        "steps": steps, "what": "reference gsplat CUDA kernels (oracle/_ref, sm_100a, the reference's "
                        "flags) driven as scene/beta_model.py:660-711 drives them, same GPU, same scene and cameras; "
                        "ratio = this library's 1-GPU-equivalent value / theirs"}
            torch.cuda.empty_cache()
    except Exception as e:  # the reference arm must never take the headline numbers down with it
        ref_line = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    drop_line = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ref_caller

        model = ref_caller.BetaModelCaller(scene_d, bg_d, requires_grad=True)
        vcs = [ref_caller.ViewpointCamera(c) for c in cams_d]

        def d_fwd(k):
            with torch.no_grad():
                model.render(vcs[k % len(vcs)])

        def d_train(k):
            out = model.render(vcs[k % len(vcs)])
            (out["render"] * v_img).sum().backward()
            for t in model.leaves():
                t.grad = None

        s0 = dropin.stats()
        ms_f = _time_ms(d_fwd, steps, warmup)
        ms_t = _time_ms(d_train, steps, warmup)
        s1 = dropin.stats()
        drop_line = {"fwd_ms": ms_f, "fwdbwd_ms": ms_t, "fused_calls": s1["fused"] - s0["fused"],
                     "fallback_calls": s1["fallback"] - s0["fallback"],
                     "what": "BetaModel.render's statements (tests/ref_caller.py) + sum-loss backward to the seven "
                             "leaf tensors through the gsplat import shim; includes torch's activations / cat / "
                             "[mask] gathers of the caller and autograd's accumulation into the leaves"}
        del model
        torch.cuda.empty_cache()
    except Exception as e:
        drop_line = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return ref_line, drop_line


# ----------------------------------------------------------------------------------------------------------------
# this library
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist

    from ubs_b200 import _lib, fused, parallel

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    # pinned host buffers (the e2e path) should sit on the NUMA node of this rank's GPU: bind before anything is pinned
    from ubs_b200 import hostmem

    affinity_before = os.sched_getaffinity(0)
    numa = hostmem.bind_to_gpu_node(physical_gpu_index(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scene, cams, bg, cfg, desc = make_workload(args.workload)
    W, H, D, N = cfg["width"], cfg["height"], scene.D, scene.N
    P = W * H
    rec = fused.pack_records(D, *[t.to(dev) for t in scene.tensors()])
    V = torch.stack([c.viewmat for c in cams]).to(dev)
    K = torch.stack([c.K for c in cams]).to(dev)
    Cp = torch.stack([c.cam_pos for c in cams]).to(dev)
    Ts = torch.tensor([c.timestamp for c in cams], device=dev) if D == 7 else None
    bgd = bg[None].to(dev)
    rz = fused.FusedRasterizer(D, N, W, H, n_cams=1, device=dev)
    rz.enable_stage_timing(False)

    def cam_of(step):
        return (step * world + rank) % N_RING

    def render(step):
        c = cam_of(step)
        return rz.forward(rec, V[c:c + 1], K[c:c + 1], Cp[c:c + 1], None if Ts is None else Ts[c:c + 1], bgd)

    g = torch.Generator(device=dev).manual_seed(1 + rank)
    v_rc = torch.randn(1, H, W, 3, device=dev, generator=g) / P
    v_ra = torch.zeros(1, H, W, 1, device=dev)
    v_rec = torch.empty_like(rec)
    trainer = parallel.DataParallelTrainer(rz, world)

    def train(step):
        c = cam_of(step)
        trainer.step(rec, V[c:c + 1], K[c:c + 1], Cp[c:c + 1], None if Ts is None else Ts[c:c + 1], bgd, v_rc, v_ra,
                     v_rec)

    sampler = ClockSampler(physical_gpu_index(local_rank))  # runs for the whole session; regions are cut out by time
    sampler.start()

    def timed(fn, steps, warmup, with_stages, finish=None):
        for w in range(warmup):
            fn(w)
        if finish is not None:
            finish()
        barrier()
        rz.enable_stage_timing(with_stages)
        n0 = lib.ubs_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            fn(warmup + k)
        if finish is not None:
            finish()  # orders the current stream (and so the closing event) behind work queued on other streams
        e1.record()
        barrier()
        clocks = sampler.region(t0, time.perf_counter())
        launches = lib.ubs_launch_count() - n0
        ms = e0.elapsed_time(e1)
        stages = rz.stage_times_ms() if with_stages else {}
        rz.enable_stage_timing(False)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, clocks, stages

    # ---- value: frames/s, inputs resident in HBM --------------------------------------------------------------
    # (a) one frame at a time on one stream, with CUDA events around every stage: the per-kernel durations of the roofline
    ms_single, _, _, stages = timed(render, args.steps, args.warmup, True)
    # (b) the headline: the same frames through fused.RenderQueue -- two rasterisers on two streams, so that one frame's
    #     compositing tail overlaps the next frame's projection / tile binning (frames are independent units)
    rz2 = fused.FusedRasterizer(D, N, W, H, n_cams=1, device=dev)
    queue = fused.RenderQueue(iter([rz, rz2]).__next__, depth=2)

    def render_queued(step):
        c = cam_of(step)
        queue.render(rec, V[c:c + 1], K[c:c + 1], Cp[c:c + 1], None if Ts is None else Ts[c:c + 1], bgd,
                     screen_space=False)  # a rendered frame needs no `meta` / backward arrays

    ms_render, launches, clocks, _ = timed(render_queued, args.steps, args.warmup, False, finish=queue.join)
    fps = world * args.steps / (ms_render / 1e3)
    # work counters over a few cameras of the ring (diagnostic kernel, outside the timed region)
    counts = None
    for k in range(0, min(args.steps, 8)):
        render(args.warmup + k)
        c = rz.work_counts()
        counts = c if counts is None else {kk: counts[kk] + c[kk] for kk in c}
    n_cnt = min(args.steps, 8)
    counts = {kk: v / n_cnt for kk, v in counts.items()}

    # ---- train it/s: forward + backward (+ NCCL allreduce of the gradient records when world > 1) ---------------
    ms_train, launches_train, clocks_train, stages_train = timed(train, args.steps, args.warmup, True)
    its = world * args.steps / (ms_train / 1e3)

    # ---- full train step: forward + L1/SSIM loss + backward (+ all-reduce) + Adam ---------------
    from ubs_b200 import training

    rec_train = rec.clone()  # Adam moves the parameters: keep the render workload's records untouched
    gt_img = torch.rand(1, 3, H, W, device=dev, generator=g)
    sharded = None
    if world > 1:
        # rows sharded over the ranks: gradient tiles go over NVLink into the owner's staging buffer from inside the
        # projection-backward kernel, the owner reduces + applies Adam and stores the new rows into every rank
        try:
            sharded = parallel.ShardedState.create(D, N)
        except Exception as e:  # symmetric memory unavailable on this box: the NCCL path still measures the step
            sys.stderr.write("bench.py: sharded step unavailable (%s); using the NCCL all-reduce path\n" % (e,))
            sharded = None
        ok = torch.tensor([1 if sharded is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            sharded = None
        if sharded is not None:
            sharded.records.copy_(rec)
            del rec_train
            rec_train = sharded.records
    tstep = training.TrainStep(rz, training.PackedAdam(D, N, device=dev, allocate_moments=sharded is None), world=world,
                               sharded=sharded)

    def train_full(step):
        c = cam_of(step)
        tstep.step(rec_train, V[c:c + 1], K[c:c + 1], Cp[c:c + 1], None if Ts is None else Ts[c:c + 1], bgd, gt_img,
                   opacity_reg=0.01, scale_reg=0.01, batch_size=world)

    ms_full, launches_full, clocks_full, stages_full = timed(train_full, args.steps, args.warmup, True)
    its_full = world * args.steps / (ms_full / 1e3)
    used_sharded = sharded is not None
    scatter_full = None
    if sharded is not None and sharded.exchange == "pull":
        # the earlier form of the sharded step, for comparison: 144/176-byte gradient-record tiles pushed into the
        # owners' staging buffers, owner-side reduce + Adam + parameter stores
        sharded.exchange = "scatter"
        ms_s, _, _, st_s = timed(train_full, args.steps, args.warmup, True)
        scatter_full = {"value": world * args.steps / (ms_s / 1e3), "unit": "it/s", "ms_per_step": ms_s / args.steps,
                        "stages_ms": {k: v[1] for k, v in st_s.items()}}
        sharded.exchange = "pull"
    del rec_train, tstep
    nccl_full = None
    paths_check = None
    if world > 1:
        # the same step through NCCL: chunk-pipelined all-reduce of the gradient records + Adam on every rank
        rec_train = rec.clone()
        tstep = training.TrainStep(rz, training.PackedAdam(D, N, device=dev), world=world)
        ms_n, _, _, st_n = timed(train_full, args.steps, args.warmup, True)
        nccl_full = {"value": world * args.steps / (ms_n / 1e3), "unit": "it/s", "ms_per_step": ms_n / args.steps,
                     "stages_ms": {k: v[1] for k, v in st_n.items()}}
        del rec_train, tstep
        if sharded is not None:
            # correctness of the default multi-GPU step, in the driver-visible line: three iterations of the sharded
            # step and of the NCCL step from identical records and cameras; the parameters must agree (the two sum the
            # ranks' gradients in different orders, so not bit for bit) and every rank must hold the same parameters
            sharded.records.copy_(rec)
            sharded.exp_avg.zero_()
            sharded.exp_avg_sq.zero_()
            sharded.barrier()
            rec_n = rec.clone()
            t_s = training.TrainStep(rz, training.PackedAdam(D, N, device=dev, allocate_moments=False), world=world,
                                     sharded=sharded)
            t_n = training.TrainStep(rz, training.PackedAdam(D, N, device=dev), world=world)
            for k in range(3):
                c = cam_of(k)
                a = (V[c:c + 1], K[c:c + 1], Cp[c:c + 1], None if Ts is None else Ts[c:c + 1], bgd, gt_img)
                t_s.step(sharded.records, *a, opacity_reg=0.01, scale_reg=0.01, batch_size=world)
                t_n.step(rec_n, *a, opacity_reg=0.01, scale_reg=0.01, batch_size=world)
            barrier()
            d = (sharded.records - rec_n).abs()
            moved = (rec_n - rec).abs().max()
            stats = torch.stack([d.max(), (d > 1e-4).float().mean(), moved]).double()
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            chk = torch.stack([sharded.records.double().sum(), rec_n.double().sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            paths_check = {"iterations": 3, "sharded_vs_nccl_max_abs": float(stats[0]),
                           "frac_entries_off_by_1e-4": float(stats[1]), "max_abs_parameter_change": float(stats[2]),
                           "all_ranks_equal_sharded": bool(lo[0] == hi[0]), "all_ranks_equal_nccl": bool(lo[1] == hi[1])}
            del rec_n, t_s, t_n
    del sharded

    # ---- e2e: host buffers in, host image out, through the public API -----------------------------------------
    h_cam = torch.empty((N_RING, 16 + 9 + 3 + 1), dtype=torch.float32).pin_memory()
    h_cam[:, :16] = torch.stack([c.viewmat for c in cams]).reshape(N_RING, 16)
    h_cam[:, 16:25] = torch.stack([c.K for c in cams]).reshape(N_RING, 9)
    h_cam[:, 25:28] = torch.stack([c.cam_pos for c in cams])
    h_cam[:, 28] = torch.tensor([c.timestamp for c in cams])
    pipe = fused.HostPipeline([rz, rz2], depth=3)

    def e2e(step):
        pipe.render_to_host(rec, h_cam[cam_of(step)], bgd)

    for w in range(args.warmup):
        e2e(w)
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e(args.warmup + k)
    pipe.drain()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = t.item()
    e2e_fps = world * args.steps / t_e2e
    h2d = h_cam[0].numel() * 4
    d2h = P * 3 * 4  # the FP32 RGB image (what BetaModel.view returns to the host)

    try:
        os.sched_setaffinity(0, affinity_before)  # the CPU baseline below uses every host core
    except Exception:
        pass
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (compositing forward) + the HBM-bound stages -------------------------
    hbm_peak, peak_src = measured_peaks()
    sm_count = lib.ubs_device_sm_count()
    clk = (clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965) * 1e6
    fp32_peak = sm_count * 128 * clk / 1e12  # T lane-instructions/s at the clock observed during the run
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    rast_ms = stages.get("rasterize_fwd", (0, float("nan")))[1]
    slots = 9.0 * counts["E_cull"] + 11.0 * counts["E_acc"]
    slots_ref = 9.0 * counts["E_test"] + 11.0 * counts["E_acc"]
    achieved = slots / (rast_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "rasterize_fwd_kernel<3>", "bound": "fp32", "achieved": achieved, "peak": fp32_peak,
        "unit": "Tinstr/s", "frac": achieved / fp32_peak, "traffic": traffic.get("rasterize_fwd_kernel"),
        "ms_per_launch": rast_ms,
        "achieved_reference_algorithm": slots_ref / (rast_ms * 1e-3) / 1e12,
        "note": "FP32 issue-slot roofline (no dense contraction -> no tensor cores): slots = 9*E_cull + 11*E_acc "
                "(SURVEY 8(d) per-evaluation costs x evaluations left after the sub-tile cull); peak = SMs x 128 "
                "lanes x observed SM clock; achieved_reference_algorithm uses E_test (no cull) instead",
    }
    vis, pairs = counts["visible"], counts["pairs"]
    rec_b = rec.shape[1] * 4
    key_bits = 32 + max(1, (rz.tw * rz.th).bit_length()) + 1
    passes = (key_bits + 7) // 8
    stage_roof = {}

    stage_kernels = {  # stage -> the kernels it launches (keys of profiles/traffic.json)
        "fused_project_fwd": ["fused_project_fwd_kernel"],
        "isect_emit_sort_offsets": ["bin_scan_kernel", "bin_emit_kernel", "segment_sort_kernel",
                                    "segment_sort_long_kernel"],
        "fused_project_bwd": ["fused_project_bwd_kernel"], "adam_step": ["adam_kernel"],
        "fused_project_bwd_adam": ["fused_project_bwd_adam_kernel"],
        "l1_ssim_loss": ["ssim_fwd_kernel", "loss_finalize_kernel", "ssim_bwd_kernel"],
    }

    def stage_traffic(name):
        ks = stage_kernels.get(name, [])
        return sum(traffic[k] for k in ks) if ks and all(k in traffic for k in ks) else None

    def hbm_stage(name, ms, nbytes):
        if ms == ms and ms > 0:
            a = nbytes / (ms * 1e-3) / 1e9
            stage_roof[name] = {"bound": "hbm", "ms": ms, "algorithmic_bytes": nbytes, "achieved": a, "peak": hbm_peak,
                                "unit": "GB/s", "frac": a / hbm_peak, "traffic": stage_traffic(name)}

    hbm_stage("fused_project_fwd", stages.get("fused_project_fwd", (0, float("nan")))[1], N * rec_b + vis * 36)
    # SURVEY 8(d) unit figure: emit 12 B/pair + onesweep I (8 + 24 p) + offsets 8 B/pair.  The tile-binning route
    # this library runs (csrc/bin_sort.cu) moves 28 B/pair (8 written, 8 read, 12 written) and is bound by the
    # ranking instructions of its per-tile shared-memory sort, not by HBM; both figures are reported.
    sort_ms = stages.get("isect_emit_sort_offsets", (0, float("nan")))[1]
    if rz.sort_mode == "bin":
        # bytes this route moves: 8 B/pair written by the emit, 8 read + 12 written by the per-tile sort.  It is NOT
        # HBM bound (the 50 MB of pairs stay in L2): the emit is bound by L2 atomic throughput, the per-tile sort by
        # instruction issue and dependent shared-memory latency (ncu) -- frac is bytes moved / time / peak
        hbm_stage("isect_emit_sort_offsets", sort_ms, pairs * 28)
        if "isect_emit_sort_offsets" in stage_roof:
            stage_roof["isect_emit_sort_offsets"].update(
                bound="l2-atomics + issue (not hbm)", route="bin",
                survey_unit_bytes_onesweep=pairs * (12 + 8 + 24 * passes + 8),
                note="frac = bytes actually moved / time / HBM peak; the SURVEY 8(d) unit (emit 12 + onesweep "
                     "8 + 24 p + offsets 8 B/pair) describes the onesweep route, which this route replaces")
    else:
        hbm_stage("isect_emit_sort_offsets", sort_ms, pairs * (12 + 8 + 24 * passes + 8))
        if "isect_emit_sort_offsets" in stage_roof:
            stage_roof["isect_emit_sort_offsets"]["route"] = rz.sort_mode
    hbm_stage("fused_project_bwd", stages_train.get("fused_project_bwd", (0, float("nan")))[1],
              2 * N * rec_b + vis * 76)
    stage_roof["rasterize_bwd"] = {"bound": "fp32", "ms": stages_train.get("rasterize_bwd", (0, float("nan")))[1],
                                   "slots": 9.0 * counts["E_cull"] + 45.0 * counts["E_acc"],
                                   "traffic": traffic.get("rasterize_bwd3_pairlane_kernel"),
                                   "kernel": "rasterize_bwd3_pairlane_kernel<4, 3, rows>"}
    hbm_stage("adam_step", stages_full.get("adam_step", (0, float("nan")))[1], 7 * N * rec_b)
    # projection backward with the Adam epilogue: read params + 2 moments, write them back, + the screen-space gradients
    hbm_stage("fused_project_bwd_adam", stages_full.get("fused_project_bwd_adam", (0, float("nan")))[1],
              6 * N * rec_b + vis * 76)
    # loss: read img + gt, write 3 maps (pass 1); read 3 maps + img + gt, write v_img (pass 2): 10 image-sized streams
    hbm_stage("l1_ssim_loss", stages_full.get("l1_ssim_loss", (0, float("nan")))[1], 10 * P * 3 * 4)
    rb = stage_roof["rasterize_bwd"]
    if rb["ms"] == rb["ms"] and rb["ms"] > 0:
        rb["achieved"] = rb["slots"] / (rb["ms"] * 1e-3) / 1e12
        rb["peak"], rb["unit"], rb["frac"] = fp32_peak, "Tinstr/s", rb["slots"] / (rb["ms"] * 1e-3) / 1e12 / fp32_peak

    # ---- the number to beat: the reference's own CUDA kernels on this GPU, and the zero-edit drop-in -------------
    ref_cuda_line, dropin_line = measure_reference_cuda_and_dropin(scene, cams, bg, cfg, dev, its / world, fps / world)

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample --------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_oracle_pass(scene, cams[0], bg, False)  # warm-up (builds / loads the C oracle, faults the pages)
        t_f = [cpu_oracle_pass(scene, cams[k + 1], bg, False) for k in range(2)]
        t_b = cpu_oracle_pass(scene, cams[3], bg, True)
        cpu = {"value": len(t_f) / sum(t_f), "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "2 whole forward frames + 1 forward+backward frame of the same workload (oracle/, "
                         "torch + C/OpenMP on all host cores)",
               "train_it_per_s": 1.0 / t_b}

    line = {
        "metric": "render_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_render / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "N": N, "D": D, "width": W, "height": H,
                   "cameras_per_gpu_per_step": 1, "parallelism": "camera-parallel x%d, no collective" % world,
                   "l2": "inputs larger than L2 (%.0f MB parameter records re-read every step)" % (N * rec_b / 1e6),
                   "pairs_per_frame": pairs, "visible_per_frame": vis, "peak_source": peak_src,
                   "sort_route": rz.sort_mode,
                   "frames_in_flight": "2 (fused.RenderQueue: two rasterisers on two streams); one frame at a time on "
                                       "one stream: %.1f frames/s, %.4f ms/frame -- the stage and roofline times are "
                                       "from that pass" % (world * args.steps / (ms_single / 1e3), ms_single / args.steps)},
        "train": {"metric": "train_it_per_s", "value": its, "unit": "it/s", "ms_per_step": ms_train / args.steps,
                  "what": "forward + backward to the packed parameter-gradient records" +
                          (" + NCCL allreduce(sum) of the %.0f MB gradient buffer" % (N * rec_b / 1e6)
                           if world > 1 else ""),
                  "gpu_launches": launches_train, "clocks": clocks_train,
                  "stages_ms": {k: v[1] for k, v in stages_train.items()}},
        "train_full": {"metric": "full_train_step_it_per_s", "value": its_full, "unit": "it/s",
                       "ms_per_step": ms_full / args.steps,
                       "what": "forward + fused L1/SSIM loss and image gradient + backward" +
                               (" + NCCL allreduce" if world > 1 else "") +
                               " + Adam over the packed records with the opacity/scale regularisers (inside the "
                               "projection-backward kernel at 1 GPU, a separate pass after the all-reduce otherwise) "
                               "(train.py:100-171 for one view per GPU)",
                       "gpu_launches": launches_full, "clocks": clocks_full,
                       "stages_ms": {k: v[1] for k, v in stages_full.items()},
                       "update": ("Adam inside the projection-backward kernel" if world == 1 else
                                  "chunk-pipelined NCCL all-reduce + Adam on every rank" if not used_sharded else
                                  "rows sharded over the ranks (symmetric memory): every rank leaves its view's 48-byte "
                                  "screen-space gradient rows in a peer-mapped buffer; the owner of a shard pulls all "
                                  "views' rows of its primitives by bulk copies over NVLink inside ONE kernel that runs "
                                  "the projection backward over the views, Adam, and stores the new rows to every rank"),
                       "nccl_allreduce_path": nccl_full, "sharded_scatter_path": scatter_full},
        "roofline": roofline, "stages": stage_roof, "stages_ms": {k: v[1] for k, v in stages.items()},
        "work": counts, "cpu_baseline": cpu,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "fused.HostPipeline.render_to_host: pinned host camera -> device, render, FP32 RGB image -> "
                        "pinned host, three frames in flight; wall clock over the timed steps",
                "host_placement": numa},
        "gpu_launches": launches, "clocks": clocks, "clocks_session": sampler.stop(),
        # ---- the last keys: what a reader of the tail of this line needs ----------------------------------------------
        "train_full_summary": {"value": its_full, "unit": "it/s", "ms_per_step": ms_full / args.steps,
                               "stages_ms": {k: round(v[1], 4) for k, v in stages_full.items()},
                               "nccl_path_it_per_s": None if nccl_full is None else nccl_full["value"],
                               "scatter_path_it_per_s": None if scatter_full is None else scatter_full["value"],
                               "paths_check": paths_check},
        "reference_cuda": ref_cuda_line, "dropin": dropin_line,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_cpu(args, rank, world)
    elif args.impl == "reference_cuda":
        run_reference_cuda(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            sys.stderr.write("bench.py: --gpus %d without torchrun: running 1 rank\n" % args.gpus)
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
