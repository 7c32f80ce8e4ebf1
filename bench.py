#!/usr/bin/env python3
"""bench.py -- interpolated frames/s of the RIFE HDv3 IFNet hot path at 1080p 2X (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|fp32] [--pairs P]
    python bench.py --impl reference ...      # the reference's CPU path (torch-CPU port, oracle/torch_ref.py)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one Model.inference call on a batch of P synthetic 1088x1920 (padded 1080p) frame pairs at
timestep 0.5 per GPU (the 2X interpolation of P pairs).  Frame pairs are independent, so ranks shard pairs with
no data-path collective (scaling = weak); torch.distributed is only used for the barrier and the max-over-ranks.
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

H, W, CH = 1088, 1920, 3          # 1080p padded to a multiple of 64 (cfg3 shape)
WORKLOAD = "1080p 2X: 1088x1920 (padded 1080p) frame pairs, timestep 0.5, scale 1.0, HDv3 v4.25 weights = seeded default init"
BYTES_PER_PIXEL_MODEL = 601.9     # stage-fused HBM traffic model, elements per pixel per pair
FLOP_PER_PIXEL = 6064 + 38372.3   # SURVEY 8d at scale = 1


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


def cpu_reference_fps(n_pairs, warmup=1):
    """The reference's CPU path (fp32, all host threads): oracle/torch_ref.py issues the same ATen operator sequence
    as train_log/RIFE_HDv3.py:56-60 (the unmodified reference only exists in the build container)."""
    import torch
    from oracle import torch_ref as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = seeded_state_dict()
    a, b = smooth_pairs(1, H, W, 0, torch.float32)
    for _ in range(warmup):
        T.model_inference(sd, a, b)
    t0 = time.perf_counter()
    for _ in range(n_pairs):
        T.model_inference(sd, a, b)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt / n_pairs, cores


def run_reference(args, rank):
    if rank != 0:
        return
    fps, sec, cores = cpu_reference_fps(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "interpolated frames/sec at 1080p 2X", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": 1, "device": "cpu"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} single 1088x1920 pairs after {args.warmup} warm-up, torch-CPU fp32 port "
                                   f"of the reference op sequence (oracle/torch_ref.py), {cores} threads"},
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
    ap.add_argument("--precision", default=os.environ.get("RIFE_BENCH_PRECISION", "fp16"), choices=["fp16", "fp32"])
    ap.add_argument("--pairs", type=int, default=16, help="frame pairs per step per GPU (batch of one Model.inference)")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("RIFE_CHUNK_PAIRS", "0")),
                    help="pairs per pass of the 1/2- and full-resolution stages (rife_plan_set_chunk); 0 = library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    pkg.model.device = dev
    torch.manual_seed(1234)
    model = pkg.Model(precision=args.precision)
    model.eval()
    net = model.flownet
    P = args.pairs
    dt = torch.float16 if args.precision == "fp16" else torch.float32
    esz = 2 if args.precision == "fp16" else 4

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

    # ---- (1) device-resident throughput: `value`
    for i in range(args.warmup):
        a, b = dev_sets[i % NSETS]
        out = model.inference(a, b)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = lib.rife_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        a, b = dev_sets[i % NSETS]
        out = model.inference(a, b)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.rife_launch_count() - n0
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = world * P * args.steps / (ms / 1e3)

    # ---- (2) end to end through the host-buffer C-ABI call (rife_inference_host): pinned host frames exactly as a video
    #      decoder delivers them -- uint8 HWC, true 1080 x 1920 (the library pads to 1088 and crops back) -> H2D ->
    #      inference -> D2H of the uint8 result, every step.  Three plan slots rotate over one H2D, one compute and one
    #      D2H stream (rife_inference_host_pipelined): the PCIe copies of a step overlap the kernels of its neighbours
    #      and the kernels of consecutive steps run back to back.
    NSLOT = 3
    FH, FW = 1080, 1920
    u8_sets = []
    for a, b in host_sets:
        fa = (a[:, :, :FH, :FW].float() * 255).byte().permute(0, 2, 3, 1).contiguous().pin_memory()
        fb = (b[:, :, :FH, :FW].float() * 255).byte().permute(0, 2, 3, 1).contiguous().pin_memory()
        u8_sets.append((fa, fb))
    host_out = [torch.empty((P, FH, FW, CH), dtype=torch.uint8).pin_memory() for _ in range(NSLOT)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSLOT)]
    cur = torch.cuda.current_stream()

    def e2e_steps(n):
        for i in range(n):
            k = i % NSLOT
            a, b = u8_sets[i % NSETS]
            model.inference_host(a, b, host_out[k], slot=k, streams=streams)

    e2e_steps(2 * NSLOT)
    barrier()
    e0.record()
    for st_ in streams:
        st_.wait_stream(cur)
    e2e_steps(args.steps)
    for st_ in streams:
        cur.wait_stream(st_)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * P * args.steps / (ms_e2e / 1e3)
    frame_bytes = P * CH * FH * FW

    # ---- (2b) the same PCIe traffic with no kernels: what the host <-> device link alone sustains with every rank
    #      copying at once (explains e2e when several GPUs share the host's PCIe / memory bandwidth)
    d_in = [torch.empty_like(u8_sets[0][0], device=dev) for _ in range(2)]
    d_out = torch.empty((P, FH, FW, CH), dtype=torch.uint8, device=dev)

    def copy_steps(n):
        for i in range(n):
            a, b = u8_sets[i % NSETS]
            with torch.cuda.stream(streams[0]):
                d_in[0].copy_(a, non_blocking=True)
                d_in[1].copy_(b, non_blocking=True)
            with torch.cuda.stream(streams[2]):
                host_out[i % NSLOT].copy_(d_out, non_blocking=True)

    copy_steps(3)
    barrier()
    e0.record()
    for st_ in streams:
        st_.wait_stream(cur)
    copy_steps(args.steps)
    for st_ in streams:
        cur.wait_stream(st_)
    e1.record()
    barrier()
    ms_copy = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    # ---- (3) per-kernel CUDA-event timing (separate pass over the same steps so that `value` is undisturbed)
    plan = net._plans[next(k for k in net._plans if k[-1] == 0)]
    lib.rife_plan_set_frame_size(plan.handle, H, W)   # slot 0 was last used by the 1080-row e2e path
    plan.frame = (H, W)
    lib.rife_plan_set_profiling(plan.handle, 1)
    acc, cnt = {}, {}
    names = (C.c_char_p * 128)()
    msbuf = (C.c_float * 128)()
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        a, b = dev_sets[i % NSETS]
        model.inference(a, b)
        torch.cuda.synchronize()
        n = lib.rife_plan_get_profile(plan.handle, names, msbuf, 128)
        for k in range(n):
            nm = names[k].decode()
            acc[nm] = acc.get(nm, 0.0) + msbuf[k]
            cnt[nm] = cnt.get(nm, 0) + 1
    lib.rife_plan_set_profiling(plan.handle, 0)
    stage_ms = {k: acc[k] / prof_steps for k in acc}          # ms per step, summed over that stage's launches
    stage_launches = {k: cnt[k] / prof_steps for k in cnt}

    hbm_peak, tf_peak, peak_src = peaks()
    N = P * H * W
    # algorithmic bytes of the dominant HBM-bound kernel, the block-4 input gather:
    #   reads src0+src1 (2 x 8 ch) + tmp_3 (16 ch at 1/4 of the pixels) + tmp_0..2 (negligible), writes 32 ch
    alg = {
        # src0 + src1 (8 ch each) + tmp_3 (16 ch at 1/4 of the pixels) + xin (32 ch), in E-byte elements,
        # plus the fp32 flow buffer read and re-written (2 x 16 bytes per pixel)
        "b4.input": N * esz * (16 + 16 / 4 + 32) + N * 32,
        "merge": N * esz * (16 + 16 + 16 / 4 + 16 / 16 + 16 / 64 + 16 / 256) + N * CH * esz,
        "prep": 2 * N * CH * esz + 2 * N * esz * (8 + 16 / 4),
    }
    top = max(stage_ms, key=lambda k: stage_ms[k])
    rk = "b4.input"
    n_l = max(stage_launches.get(rk, 1), 1)            # launches of that kernel per step (= passes, --chunk)
    k_ms = stage_ms.get(rk, 0.0) / n_l
    alg_launch = alg[rk] / n_l
    achieved = alg_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None   # dram bytes per launch of that kernel from the committed ncu --set full capture (same workload only)
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp) and args.precision == "fp16":
        tj = json.load(open(tp))
        traffic = tj.get("traffic_bytes_per_launch") * (P / n_l) / float(tj.get("pairs", 4))   # bytes scale with the pairs per launch
    roofline = {"bound": "hbm", "kernel": rk, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_launch, "kernel_ms": k_ms, "launches_per_step": n_l,
                "top_stage_by_time": top,
                "pipeline_model": {"bytes_per_pair_stage_fused_model": BYTES_PER_PIXEL_MODEL * H * W * esz,
                                   "frac_of_hbm_peak": (BYTES_PER_PIXEL_MODEL * H * W * esz * value / world) / (hbm_peak * 1e9),
                                   "tflops": FLOP_PER_PIXEL * H * W * value / world / 1e12},
                "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}}

    line = {
        "metric": "interpolated frames/sec at 1080p 2X", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P, "pairs_per_pass_full_res_stages": int(P / n_l + 0.5),
                   "precision": args.precision,
                   "parallelism": f"frame-pair sharding x{world}, no collective",
                   "host_numa_cores_bound": numa_cores,
                   "l2_policy": f"inputs rotate over {NSETS} distinct batches; each step streams >1 GB of intermediates (> 126 MB L2)"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": 2 * frame_bytes, "d2h_bytes_per_step": frame_bytes,
                "ms_per_step": ms_e2e / args.steps, "copy_only_ms_per_step": ms_copy,
                "api": "rife_inference_host: pinned uint8 HWC 1080x1920 frames in, uint8 HWC frame out, 3 plan slots over h2d / compute / d2h streams (rife_inference_host_pipelined)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, sec, cores = cpu_reference_fps(6, 1)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"6 single 1088x1920 pairs after 1 warm-up ({sec:.2f} s each), torch-CPU fp32 port of "
                                          f"the reference op sequence (oracle/torch_ref.py), {cores} threads"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
