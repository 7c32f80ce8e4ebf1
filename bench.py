#!/usr/bin/env python3
"""bench.py -- interpolated frames/s of the RIFE HDv3 IFNet hot path at 1080p 2X (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|bf16|fp32] [--pairs P] [--workload NAME]
    python bench.py --impl reference ...      # the reference's CPU path (the unmodified reference when RIFE_REFERENCE /
                                              # baseline/_ref holds it, else the torch-CPU port oracle/torch_ref.py)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one Model.inference call on a batch of P synthetic 1088x1920 (padded 1080p) frame pairs at
timestep 0.5 per GPU (the 2X interpolation of P pairs).  Frame pairs are independent, so ranks shard pairs with
no data-path collective (scaling = weak); torch.distributed is only used for the barrier and the max-over-ranks.
--workload selects the other BASELINE.json configs (720p2x = config 2, 1080p4x = config 3, 4k2x = config 4).
Prints ONE JSON line on rank 0 (see the contract in the task description).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CH = 3
# BASELINE.json configs (cfg2 / cfg3 / cfg4 shapes): padded size, true frame size, pairs per step, timesteps, scale
WORKLOADS = {
    "1080p2x": dict(H=1088, W=1920, fh=1080, fw=1920, pairs=16, ts=(0.5,), scale=1.0,
                    text="1080p 2X: 1088x1920 (padded 1080p) frame pairs, timestep 0.5, scale 1.0"),
    "720p2x": dict(H=768, W=1280, fh=720, fw=1280, pairs=32, ts=(0.5,), scale=1.0,
                   text="720p 2X: 768x1280 (padded 720p) frame pairs, timestep 0.5, scale 1.0"),
    "1080p4x": dict(H=1088, W=1920, fh=1080, fw=1920, pairs=6, ts=(0.25, 0.5, 0.75), scale=1.0,
                    text="1080p 4X: 1088x1920 frame pairs at timesteps 0.25 / 0.5 / 0.75 in one call (Head once per frame), scale 1.0"),
    "4k2x": dict(H=2176, W=3840, fh=2160, fw=3840, pairs=4, ts=(0.5,), scale=0.5,
                 text="4K 2X: 2176x3840 (padded 2160p) frame pairs, timestep 0.5, scale 0.5 (scale_list [32,16,8,4,2])"),
}
WEIGHTS_TEXT = "HDv3 v4.25 weights = seeded default init"
BYTES_PER_PIXEL_MODEL = 601.9     # stage-fused HBM traffic model, elements per pixel per pair
FLOP_PER_PIXEL = 6064 + 38372.3   # SURVEY 8d at scale = 1
BLOCK_C = (192, 128, 96, 64, 32)
# DRAM bytes per full-resolution output pixel (dram__bytes_read.sum + dram__bytes_write.sum) of the three largest stages, from
# the committed `ncu --set full` capture of one 8-pair 1080p pass of this build (profiles/r02_ncu_full_hdv3_1080p_P8.md rows 53,
# 41, 64: 1317.2 / 1074.7 / 1166.2 MB over 8 x 1088 x 1920 pixels).  Per-pixel figures scale with the batch, not with a clock.
NCU_DRAM_BYTES_PER_PX = {"b4.input+conv0.0": 78.8, "b3.input": 64.3, "merge": 69.8}


def stage_model(P, H, W, esz, io_esz, nframes):
    """Algorithmic (compulsory) HBM bytes and real FLOPs of every pipeline stage of ONE step at scale = 1 -- each kernel
    reads its inputs once and writes its outputs once.  Keys = the stage names of
    rife_plan_get_profile; P outputs, `nframes` SRC frames (2P for independent pairs)."""
    N, NF = P * H * W, nframes * H * W
    m = {}
    m["prep"] = (NF * CH * io_esz + NF / 4 * 16 * esz, 0.0)
    m["head.cnn0"] = (2 * NF / 4 * 16 * esz, 2.0 * NF / 4 * 27 * 16)
    one = (2 * NF / 4 * 16 * esz, 2.0 * NF / 4 * 9 * 16 * 16)
    m["head.cnn1"] = m["head.cnn2"] = one
    m["head.cnn1+2"] = (one[0], 2 * one[1])                      # fused pair: the intermediate never reaches HBM
    # the four layers in one kernel (conv_head.cu): the s2d frame in, SRC out; the real FLOPs of the four layers
    m["head.fused"] = (NF / 4 * 16 * esz + NF * 8 * esz,
                       2.0 * NF / 4 * 27 * 16 + 2 * one[1] + 2.0 * NF * 4 * 4 * 16)
    m["head.cnn3"] = (NF / 4 * 16 * esz * 2 + NF * 8 * esz, 2.0 * NF * 4 * 4 * 16)   # features + s2d frame in, SRC out
    for i, c in enumerate(BLOCK_C):
        s = 16 >> i
        n = N / (s * s)
        p_in = 15 if i == 0 else 28
        if i == 0:
            gather = N * 4 / 256 * 2 * 8 * esz
        else:
            samp = N if s == 1 else N * 4 / (s * s)             # full-resolution pixels the 1/s resize reads
            gather = samp * 2 * 8 * esz                          # one 8-channel SRC pixel per frame per sample
            gather += N / ((2 * s) ** 2) * 16 * esz              # newest tmp level (block i - 1)
            if s <= 2:
                gather += samp * 16 * (2 if s == 1 else 1)       # fp32 flow buffer: written at s = 2, read + rewritten at s = 1
        m[f"b{i}.input"] = (gather + n * 32 * esz, 0.0)
        m[f"b{i}.conv0.0"] = (n * 32 * esz + n / 4 * (c // 2) * esz, 2.0 * n / 4 * 9 * p_in * (c // 2))
        if s == 1:
            # gather fused into conv0.0 (conv_gather.cu): the block input never reaches HBM; the flow buffer is read and a
            # second one written
            m[f"b{i}.input+conv0.0"] = (gather + n / 4 * (c // 2) * esz, 2.0 * n / 4 * 9 * p_in * (c // 2))
        m[f"b{i}.conv0.1"] = (n / 4 * (c // 2) * esz + n / 16 * c * esz, 2.0 * n / 16 * 9 * (c // 2) * c)
        # 8 ResConv layers; the fused pairs of block 4 (conv_strip.cu) read and write HBM once per two layers
        m[f"b{i}.res"] = (8 * 2 * n / 16 * c * esz, 8 * 2.0 * n / 16 * 9 * c * c)
        m[f"b{i}.res/fused-pairs"] = (4 * 2 * n / 16 * c * esz, 8 * 2.0 * n / 16 * 9 * c * c)
        tmp_c = 8 if i == 4 else 16                              # Model.inference: flow + mask only from the last block
        m[f"b{i}.lastconv"] = (n / 16 * c * esz + n * tmp_c * esz, 2.0 * n / 16 * 16 * c * (20 if i == 4 else 52))
    m["merge"] = (N * (16 + 8 * esz + 2 * 4 * esz + CH * io_esz), 0.0)   # flow buffer + tmp4 + image halves of 2 SRC pixels + frame out
    return m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled every 5 ms DURING the timed region through NVML (nvidia-ml-py);
    falls back to `nvidia-smi -lms 200` when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.proc = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through its UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.rows = []
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        masks = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200"}


def seeded_state_dict():
    import torch
    pkg = importlib.import_module("eccv2022-rife_b200")
    torch.manual_seed(1234)
    net = pkg.IFNet()
    return {k: v.detach().cpu().float() for k, v in net.state_dict().items()}


def smooth_pairs(n, h, w, seed, dtype):
    """Synthetic frames: band-limited noise (coherent flow) plus 5% white noise (high-frequency texture)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(n, CH, h // 16 + 2, w // 16 + 2, generator=g)
    big = torch.nn.functional.interpolate(lo, size=(h + 32, w + 32), mode="bicubic", align_corners=False)
    a = big[:, :, 16:16 + h, 16:16 + w] + 0.05 * torch.rand(n, CH, h, w, generator=g)
    b = big[:, :, 12:12 + h, 22:22 + w] + 0.05 * torch.rand(n, CH, h, w, generator=g)
    return a.clamp(0, 1).to(dtype).contiguous(), b.clamp(0, 1).to(dtype).contiguous()


def find_reference():
    """The unmodified reference, when this box has it: $RIFE_REFERENCE, baseline/_ref, or the build container's mount."""
    for d in (os.environ.get("RIFE_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.exists(os.path.join(d, "train_log", "RIFE_HDv3.py")):
            return d
    return None


def cpu_reference_fps(wl, n_calls, warmup=1):
    """The reference's CPU path (fp32, all host threads) on single frame pairs of the workload's shape.  kind "reference":
    the unmodified train_log/RIFE_HDv3.Model imported from find_reference(), same seeded weights; kind "port":
    oracle/torch_ref.py issues the same ATen operator sequence as train_log/RIFE_HDv3.py:56-60 (the reference is pure
    Python on PyTorch, has no setup.py to install from, and only exists in the build container)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_grad_enabled(False)
    sd = seeded_state_dict()
    a, b = smooth_pairs(1, wl["H"], wl["W"], 0, torch.float32)
    ref_dir = find_reference()
    kind = "port"
    if ref_dir:
        try:
            sys.dont_write_bytecode = True
            sys.path.insert(0, ref_dir)
            from train_log.RIFE_HDv3 import Model as RefModel   # the reference itself
            import model.warplayer as ref_wl
            ref_wl.device = torch.device("cpu")
            m = RefModel()
            m.flownet.load_state_dict(sd)
            m.flownet.to("cpu")
            m.eval()
            run = lambda t: m.inference(a, b, t, wl["scale"])
            kind = "reference"
        except Exception as e:   # missing optional dependency of the reference's import chain etc.
            sys.stderr.write(f"bench: unmodified reference at {ref_dir} not importable ({e}); using the port\n")
            sys.path.remove(ref_dir)
    if kind == "port":
        from oracle import torch_ref as T
        run = lambda t: T.model_inference(sd, a, b, timestep=t, scale=wl["scale"])
    ts = wl["ts"]
    for i in range(warmup):
        run(ts[i % len(ts)])
    t0 = time.perf_counter()
    for i in range(n_calls):
        run(ts[i % len(ts)])
    dt = time.perf_counter() - t0
    return n_calls / dt, dt / n_calls, cores, kind


def metric_name(wname):
    return "interpolated frames/sec at 1080p 2X" if wname == "1080p2x" else f"interpolated frames/sec, workload {wname}"


def run_reference(args, rank):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    fps, sec, cores, kind = cpu_reference_fps(wl, args.steps, args.warmup)
    what = ("unmodified reference train_log/RIFE_HDv3.Model.inference on CPU" if kind == "reference" else
            "torch-CPU fp32 port of the reference op sequence (oracle/torch_ref.py)")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["text"] + ", " + WEIGHTS_TEXT, "pairs_per_step": 1, "device": "cpu"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} single {wl['H']}x{wl['W']} pairs after {args.warmup} warm-up, {what}, {cores} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RIFE_BENCH_PRECISION", "fp16"), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--workload", default="1080p2x", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per step per GPU (0 = the workload's default)")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("RIFE_CHUNK_PAIRS", "0")),
                    help="pairs per pass of the 1/2- and full-resolution stages (rife_plan_set_chunk); 0 = library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-leg", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.chunk:
        os.environ["RIFE_CHUNK_PAIRS"] = str(args.chunk)    # read once by the library
    pkg = importlib.import_module("eccv2022-rife_b200")
    lib = pkg._lib.load()
    numa_cores = pkg.sharding.bind_to_gpu_numa(local_rank) if world > 1 else 0   # pinned buffers on the GPU's NUMA node
    torch.manual_seed(1234)
    model = pkg.Model(precision=args.precision)
    model.eval()
    net = model.flownet
    wl = WORKLOADS[args.workload]
    H, W, FH, FW = wl["H"], wl["W"], wl["fh"], wl["fw"]
    P = args.pairs or wl["pairs"]
    TS, scale = wl["ts"], wl["scale"]
    NT = len(TS)
    dt = torch.float32 if args.precision == "fp32" else torch.float16
    esz = 4 if args.precision == "fp32" else 2

    # ---- synthetic inputs: NSETS distinct batches so that no step finds its inputs in L2 (each step also streams
    #      far more than the 126 MB L2 through its intermediates)
    NSETS = 3
    host_sets = [smooth_pairs(P, H, W, 100 * rank + i, dt) for i in range(NSETS)]
    host_sets = [(a.pin_memory(), b.pin_memory()) for a, b in host_sets]
    dev_sets = [(a.to(dev), b.to(dev)) for a, b in host_sets]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def infer_dev(a, b):
        if NT == 1:
            return model.inference(a, b, TS[0], scale)
        return model.inference_multi(a, b, TS, scale)

    # ---- (1) device-resident throughput: `value`
    for i in range(args.warmup):
        a, b = dev_sets[i % NSETS]
        out = infer_dev(a, b)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = lib.rife_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        a, b = dev_sets[i % NSETS]
        out = infer_dev(a, b)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.rife_launch_count() - n0
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = world * P * NT * args.steps / (ms / 1e3)

    # ---- (2) end to end through the host-buffer C-ABI call (rife_inference_host_pipelined): pinned host frames exactly
    #      as a video decoder delivers them -- uint8 HWC, true frame size (the library pads and crops) -> H2D -> inference
    #      -> D2H of the uint8 result, every step.  Three plan slots rotate over one H2D, one compute and one D2H stream:
    #      the PCIe copies of a step overlap the kernels of its neighbours.
    #      (a) `e2e`, video order (inference_video.py:215-270 walks consecutive frames; "2X" = one new frame between
    #          every two decoded ones): a step uploads P + 1 consecutive frames and downloads the P in-between frames --
    #          every decoded frame crosses the host link (and the Head) once;
    #      (b) `e2e_pairs`: P independent pairs per step (2P frames up, P down), the round-1 definition.
    NSLOT = 3
    u8 = lambda t: (t[:, :, :FH, :FW].float() * 255).byte().permute(0, 2, 3, 1).contiguous().pin_memory()
    u8_sets = [(u8(a), u8(b)) for a, b in host_sets]
    seq_sets = [torch.cat((fa, fb[:1]), 0).contiguous().pin_memory() for fa, fb in u8_sets]    # P + 1 frames
    host_out = [torch.empty((P, FH, FW, CH), dtype=torch.uint8).pin_memory() for _ in range(NSLOT)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSLOT)]
    cur = torch.cuda.current_stream()
    frame_bytes = CH * FH * FW
    e2e_legs = {}

    def timed(fn, nwarm):
        fn(nwarm)
        barrier()
        e0.record()
        for st_ in streams:
            st_.wait_stream(cur)
        fn(args.steps)
        for st_ in streams:
            cur.wait_stream(st_)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    if NT == 1:
        def seq_steps(n):
            for i in range(n):
                model.inference_host(seq_sets[i % NSETS], None, host_out[i % NSLOT], TS[0], scale, slot=i % NSLOT,
                                     streams=streams, device=dev)

        def pair_steps(n):
            for i in range(n):
                a, b = u8_sets[i % NSETS]
                model.inference_host(a, b, host_out[i % NSLOT], TS[0], scale, slot=i % NSLOT, streams=streams, device=dev)

        for name, fn, up in (("video", seq_steps, P + 1), ("pairs", pair_steps, 2 * P)):
            ms_leg = timed(fn, 2 * NSLOT)
            e2e_legs[name] = {"value": world * P * args.steps / (ms_leg / 1e3), "unit": "frames/s",
                              "h2d_bytes_per_step": up * frame_bytes, "d2h_bytes_per_step": P * frame_bytes,
                              "ms_per_step": ms_leg / args.steps}
    else:
        # several timesteps per pair: pinned uint8 frames -> device -> inference_multi -> pinned uint8 result, one stream
        multi_out = torch.empty((P, NT, FH, FW, CH), dtype=torch.uint8).pin_memory()
        d_in = [torch.empty_like(u8_sets[0][0], device=dev) for _ in range(2)]

        def multi_steps(n):
            for i in range(n):
                a, b = u8_sets[i % NSETS]
                d_in[0].copy_(a, non_blocking=True)
                d_in[1].copy_(b, non_blocking=True)
                multi_out.copy_(model.inference_multi(d_in[0], d_in[1], TS, scale, pad=True), non_blocking=True)

        ms_leg = timed(multi_steps, 3)
        e2e_legs["pairs"] = {"value": world * P * NT * args.steps / (ms_leg / 1e3), "unit": "frames/s",
                             "h2d_bytes_per_step": 2 * P * frame_bytes, "d2h_bytes_per_step": P * NT * frame_bytes,
                             "ms_per_step": ms_leg / args.steps}

    # ---- (2b) the same PCIe traffic with no kernels, each direction alone and both at once: what the host <-> device
    #      link sustains with every rank copying at the same time (names the limit of e2e on a box whose GPUs share it)
    d_seq = torch.empty_like(seq_sets[0], device=dev)
    d_out = torch.empty((P, FH, FW, CH), dtype=torch.uint8, device=dev)

    def copy_steps(h2d, d2h):
        def fn(n):
            for i in range(n):
                if h2d:
                    with torch.cuda.stream(streams[0]):
                        d_seq.copy_(seq_sets[i % NSETS], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(streams[2]):
                        host_out[i % NSLOT].copy_(d_out, non_blocking=True)
        return fn

    ms_h2d = timed(copy_steps(True, False), 3) / args.steps
    ms_d2h = timed(copy_steps(False, True), 3) / args.steps
    ms_both = timed(copy_steps(True, True), 3) / args.steps
    link = {"h2d_GBps_per_gpu": (P + 1) * frame_bytes / ms_h2d / 1e6, "d2h_GBps_per_gpu": P * frame_bytes / ms_d2h / 1e6,
            "both_ms_per_step": ms_both, "h2d_alone_ms_per_step": ms_h2d, "d2h_alone_ms_per_step": ms_d2h,
            "what": "video-order copies of one step (P + 1 frames up, P frames down), no kernels, all ranks at once"}

    # ---- (3) per-kernel CUDA-event timing (separate pass over the same steps so that `value` is undisturbed)
    plan = net._plans[next(k for k in net._plans if k[-3] == 0 and k[1] == P and not k[-1] and k[-2] == NT)]
    lib.rife_plan_set_frame_size(plan.handle, H, W)   # slot 0 was last used by the true-frame-size e2e path
    plan.frame = (H, W)
    lib.rife_plan_set_profiling(plan.handle, 1)
    acc, cnt = {}, {}
    names = (C.c_char_p * 128)()
    msbuf = (C.c_float * 128)()
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        a, b = dev_sets[i % NSETS]
        infer_dev(a, b)
        torch.cuda.synchronize()
        n = lib.rife_plan_get_profile(plan.handle, names, msbuf, 128)
        for k in range(n):
            nm = names[k].decode()
            acc[nm] = acc.get(nm, 0.0) + msbuf[k]
            cnt[nm] = cnt.get(nm, 0) + 1
    lib.rife_plan_set_profiling(plan.handle, 0)
    stage_ms = {k: acc[k] / prof_steps for k in acc}          # ms per step, summed over that stage's launches
    stage_launches = {k: cnt[k] / prof_steps for k in cnt}

    # ---- (4) roofline: every stage against its own bound.  t_hbm = algorithmic bytes / measured copy peak, t_tc = real
    #      FLOPs / measured sustained dense bf16 rate (MEASURED_PEAKS.json); the stage's bound is the larger of the two and
    #      frac = that time / the CUDA-event time of the stage.  Scale-1, single-timestep workloads only.
    hbm_peak, tf_peak, peak_src = peaks()
    stages = []
    if scale == 1.0 and NT == 1:
        sm = stage_model(P, H, W, esz, esz, 2 * P)
        for nm, t_ms in sorted(stage_ms.items(), key=lambda kv: -kv[1]):
            key = nm
            if nm == "b4.res" and stage_launches.get(nm, 8) <= 4.5:
                key = "b4.res/fused-pairs"
            if key not in sm or t_ms <= 0:
                continue
            by, fl = sm[key]
            t_hbm, t_tc = by / (hbm_peak * 1e9), fl / (tf_peak * 1e12)
            bound = "tensor" if t_tc > t_hbm else "hbm"
            ach = fl / (t_ms * 1e-3) / 1e12 if bound == "tensor" else by / (t_ms * 1e-3) / 1e9
            stages.append({"stage": nm, "ms_per_step": round(t_ms, 4), "launches_per_step": stage_launches[nm],
                           "bound": bound, "alg_bytes": by, "flops": fl, "achieved": round(ach, 1),
                           "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                           "frac": round(max(t_hbm, t_tc) / (t_ms * 1e-3), 4),
                           "hbm_frac": round(t_hbm / (t_ms * 1e-3), 4), "tensor_frac": round(t_tc / (t_ms * 1e-3), 4)})
    if stages:
        dom = stages[0]                                        # dominant stage by time
        worst = min((s_ for s_ in stages if s_["ms_per_step"] >= 0.01 * sum(stage_ms.values())), key=lambda s_: s_["frac"])
        n_l = max(dom["launches_per_step"], 1)
        roofline = {"bound": dom["bound"], "kernel": dom["stage"], "achieved": dom["achieved"],
                    "peak": tf_peak if dom["bound"] == "tensor" else hbm_peak, "unit": dom["unit"], "frac": dom["frac"],
                    "traffic": (NCU_DRAM_BYTES_PER_PX[dom["stage"]] * P * H * W / n_l
                                if dom["stage"] in NCU_DRAM_BYTES_PER_PX and (H, W) == (1088, 1920) else None),
                    "traffic_source": "ncu --set full capture of this build at 8 pairs (profiles/r02_ncu_full_hdv3_1080p_P8.md), "
                                      "DRAM read + written per output pixel x the pixels of this launch",
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom["alg_bytes"] / n_l, "kernel_ms": dom["ms_per_step"] / n_l,
                    "launches_per_step": n_l,
                    "worst_stage": {k: worst[k] for k in ("stage", "bound", "frac", "ms_per_step")},
                    "time_weighted_frac": round(sum(s_["frac"] * s_["ms_per_step"] for s_ in stages) /
                                                sum(s_["ms_per_step"] for s_ in stages), 4),
                    # whole pipeline: the algorithmic bytes of the stages as IMPLEMENTED (sum of the table above; lazy flow,
                    # no full-resolution intermediates, block 4's input never in HBM) against the HBM peak, and -- for
                    # reference only -- SURVEY 8d's stage-fused traffic MODEL of the reference's op sequence, which assumes
                    # ~2x those bytes (a rate above the HBM peak there means the implementation moves fewer bytes than
                    # the model, not that any kernel exceeds the roofline)
                    "pipeline_model": {"implemented_alg_bytes_per_pair": sum(s_["alg_bytes"] for s_ in stages) / P,
                                       "frac_of_hbm_peak": (sum(s_["alg_bytes"] for s_ in stages) / (ms_per_step * 1e-3)) / (hbm_peak * 1e9),
                                       "survey_model_bytes_per_pair": BYTES_PER_PIXEL_MODEL * H * W * esz,
                                       "survey_model_rate_over_hbm_peak": (BYTES_PER_PIXEL_MODEL * H * W * esz * value / world) / (hbm_peak * 1e9),
                                       "tflops": FLOP_PER_PIXEL * H * W * value / world / 1e12},
                    "stages": stages}
    else:
        roofline = {"bound": "hbm", "kernel": None, "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                    "traffic": None, "peak_source": peak_src,
                    "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}}

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": wl["text"] + ", " + WEIGHTS_TEXT, "workload_name": args.workload,
                   "pairs_per_step_per_gpu": P, "outputs_per_step_per_gpu": P * NT,
                   "precision": args.precision,
                   "parallelism": f"frame-pair sharding x{world}, no collective",
                   "host_numa_cores_bound": numa_cores,
                   "l2_policy": f"inputs rotate over {NSETS} distinct batches; each step streams >1 GB of intermediates (> 126 MB L2)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    api = ("Model.inference_host -> rife_inference_host_pipelined: pinned uint8 HWC %dx%d frames in, uint8 HWC frames out, "
           "3 plan slots over h2d / compute / d2h streams" % (FH, FW))
    if "video" in e2e_legs:
        line["e2e"] = dict(e2e_legs["video"], api=api + "; VIDEO ORDER: each step uploads P + 1 consecutive frames and "
                           "downloads the P in-between frames (2X interpolation of a decoded stream: every frame crosses "
                           "the link and the Head once)", host_link=link)
        line["e2e_pairs"] = dict(e2e_legs["pairs"], api=api + "; P independent frame pairs per step (round-1 definition)")
    else:
        line["e2e"] = dict(e2e_legs["pairs"], api="pinned uint8 frames -> device -> Model.inference_multi -> pinned uint8 frames, one stream",
                           host_link=link)

    # ---- (5) the fp32 mode every fp32-tensor caller of the drop-in lands on (precision="auto"), for the record
    if args.precision != "fp32" and not args.no_fp32_leg and NT == 1:
        del dev_sets, host_sets
        torch.cuda.empty_cache()
        m32 = pkg.Model(precision="fp32")
        m32.flownet.load_state_dict(net.state_dict())
        m32.eval()
        P32 = min(P, 4)
        a32, b32 = smooth_pairs(P32, H, W, 7, torch.float32)
        a32, b32 = a32.to(dev), b32.to(dev)
        for _ in range(2):
            m32.inference(a32, b32, TS[0], scale)
        barrier()
        e0.record()
        for _ in range(3):
            m32.inference(a32, b32, TS[0], scale)
        e1.record()
        barrier()
        line["fp32_value"] = {"value": world * P32 * 3 / (max_over_ranks(e0.elapsed_time(e1)) / 1e3), "unit": "frames/s",
                              "pairs_per_step_per_gpu": P32, "steps": 3,
                              "what": "the same workload in RIFE_PRECISION_FP32 (<= 1e-4 vs the reference), device-resident"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, sec, cores, kind = cpu_reference_fps(wl, 6, 1)
        what = ("unmodified reference (train_log/RIFE_HDv3.Model.inference, CPU)" if kind == "reference" else
                "torch-CPU fp32 port of the reference op sequence (oracle/torch_ref.py)")
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": f"6 single {H}x{W} pairs after 1 warm-up ({sec:.2f} s each), {what}, {cores} threads"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
