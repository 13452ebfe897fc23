#!/usr/bin/env python
"""bench.py — headline benchmark: scenes/demo2.yml at 16384 spp (800x600, depth 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--root R] [--impl reference]

A step is one render of the whole frame.  Our arm renders on N GPUs (one process per GPU
under torchrun; interleaved row tiles; one framebuffer gather over NCCL) through the C-ABI of
include/fluxb200.h.  Rank 0 prints ONE JSON line:

  value     Msamples/s, whole job, sample sets / scene resident in HBM, timed with CUDA events
            (max over ranks), kernels + gather
  e2e       same metric through the host-buffer C-ABI: every step uploads the scene, generates
            the sample sets on the device (the reference's timer also covers Camera::new,
            manager.rs:145-170) and reads the framebuffer back into host memory
  roofline  FP64-pipe roofline of the render kernel: algorithmic FP64 ops (SURVEY.md §8d model x
            device event counters) / kernel time, against the unfused FP64 issue rate measured
            live on the same GPU (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the CPU oracle (a C++/OpenMP restatement of the reference; the Rust reference
            cannot be built in this image) on all host cores, on a bounded sample

--impl reference times only that CPU oracle (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SCENE = os.path.join(ROOT, "scenes", "demo2.yml")
TILE_ROWS = 4
MAX_DEPTH = 5
SEED = 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """The host cores this process may run on.  Asked of the OS, not of OpenMP: torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would quietly turn the CPU arm into a single-threaded run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_run(root: int, threads: int = 0):
    """One full-frame render of demo2 on the CPU oracle at sample_root `root`; returns (seconds, Msamples/s)."""
    threads = threads or host_threads()
    from flux_b200 import JobConfiguration, SceneData
    from oracle import oracle_py as O
    sd = SceneData.from_yaml(SCENE)
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, MAX_DEPTH, 50)
    flat = sd.flatten()
    ss = O.generate_samples(SEED, root, MAX_DEPTH, W)
    ss.set_index = O.generate_set_index(SEED, H, W, W)
    t0 = time.perf_counter()
    O.render_rows(flat, cfg, ss, 0, H - 1, threads=threads)
    dt = time.perf_counter() - t0
    return dt, W * H * root * root / dt / 1e6


def pick_cpu_root(target_s: float) -> int:
    """Choose sample_root so that one oracle frame takes about target_s (per-sample cost is spp-independent)."""
    dt, msps = cpu_oracle_run(4)
    per_root1 = 800 * 600 / (msps * 1e6)
    root = int(max(4, min(128, (target_s / per_root1) ** 0.5)))
    return root


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_threads()
    root = args.cpu_root or pick_cpu_root(8.0)
    for _ in range(args.warmup):
        cpu_oracle_run(min(root, 8))
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_oracle_run(root)
        times.append(dt)
    per_step = sum(times) / len(times)
    msps = 800 * 600 * root * root / per_step / 1e6
    sample = (f"each step = full 800x600 demo2 frame at sample_root {root} ({root * root} spp), depth {MAX_DEPTH}; "
              f"Msamples/s is spp-independent (16384-spp frame would take {800 * 600 * 16384 / (msps * 1e6):.0f} s)")
    line = {
        "impl": "reference", "metric": "demo2.yml render throughput", "value": msps, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": msps / 5.314, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.root, args.gpus),
        "cpu_baseline": {"value": msps, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(root: int, gpus: int) -> dict:
    return {"workload": f"scenes/demo2.yml 800x600 at {root * root} spp (sample_root {root}), max_trace_depth {MAX_DEPTH}, "
                        f"13 shapes (12 spheres + 1 plane), thin-lens camera",
            "sample_sets": "800 sets x (pixel CMJ, disc CMJ, 5 hemisphere MJ), generated on device, seed 1",
            "sharding": f"interleaved tiles of {TILE_ROWS} rows over {gpus} GPU(s), one framebuffer gather",
            "l2": f"sample sets {800 * root * root * (32 + 24 * MAX_DEPTH) / 1e6:.0f} MB per GPU "
                  f"{'exceed' if 800 * root * root * (32 + 24 * MAX_DEPTH) > 126e6 else 'fit in'} the 126 MB L2; "
                  "no cross-step reuse of outputs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--root", type=int, default=128, help="sample_root (128 = 16384 spp, the headline)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-root", type=int, default=0, help="sample_root of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--kernel-mode", type=int, default=0, help="flux_set_kernel_mode (0 = auto); A/B timing only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    from flux_b200 import JobConfiguration, SceneData
    from flux_b200.opsmodel import algorithmic_ops
    from flux_b200.worker import GpuContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flux_b200 has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sd = SceneData.from_yaml(SCENE)
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    root = args.root
    cfg = JobConfiguration(root, MAX_DEPTH, 50)
    flat = sd.flatten()
    ctx = GpuContext(local_rank)
    ctx.set_kernel_mode(args.kernel_mode)
    ctx.set_scene(flat, cfg)
    ctx.generate_samples(SEED, W)
    from flux_b200.sharding import FrameGather, FramePlan
    plan = FramePlan(H, W, TILE_ROWS, world)
    my_rows = plan.my_rows(rank)
    stream = torch.cuda.current_stream().cuda_stream
    fg = FrameGather(plan, rank, dev, dist)   # packed slice (padded), gathered slices, assembled frame
    mine = fg.mine
    host_frame = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()

    def step_resident():
        ctx.render_row_list_device(my_rows, mine.data_ptr(), stream)
        fg.gather()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    total_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    # render-kernel time of the last step on this rank (CUDA events on the launching stream, inside the library)
    last_kernel_ms = ctx.last_kernel_ms()
    t = torch.tensor([total_ms, last_kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    samples_per_step = W * H * root * root
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers through the C-ABI; scene upload + sample generation + render + D2H per step
    e2e_steps = max(1, args.e2e_steps)
    out_host = np.empty((len(my_rows), W, 3), np.float64)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.set_scene(flat, cfg)
        ctx.generate_samples(SEED, W)
        if world > 1:
            ctx.render_row_list_device(my_rows, mine.data_ptr(), stream)
            frame = fg.gather()
            if rank == 0:
                host_frame.copy_(frame, non_blocking=False)
        else:
            ctx.render_row_list(my_rows, W, out=out_host)
        barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # the frame of the last e2e step, as rank 0 holds it: the same bytes whatever the number of GPUs (SURVEY.md §4:
    # sharding must not change per-pixel arithmetic), so the N = 1, 2, 4, 8 lines can be compared
    frame_sha256 = None
    if rank == 0:
        import hashlib
        final = host_frame.numpy() if world > 1 else out_host
        frame_sha256 = hashlib.sha256(np.ascontiguousarray(final).tobytes()).hexdigest()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    scene_bytes = int(flat.struct.n_spheres * (12 * 8 + 8) + flat.struct.n_planes * (6 * 8 + 8) + flat.struct.n_materials * 56
                      + 4 * len(my_rows))
    e2e = {"value": samples_per_step / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes,
           "d2h_bytes_per_step": int(H * W * 3 * 8),
           "includes": "flux_set_scene + flux_generate_samples (device) + render + framebuffer to host memory"}

    # ---- roofline: algorithmic FP64 ops of one launch / kernel time vs measured FP64 issue rate
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.render_row_list_device(my_rows, mine.data_ptr(), stream)
    torch.cuda.synchronize()
    cn = ctx.counters()
    ctx.enable_counters(False)
    ops_launch = algorithmic_ops(cn)
    peak_ginstr = ctx.measure_fp64_peak()
    achieved_tops = ops_launch / (last_kernel_ms * 1e-3) / 1e12
    # DRAM traffic per launch: dram__bytes_read + dram__bytes_write of the render kernel from the committed
    # `ncu --set full` capture (profiles/traffic.json says which), scaled by samples per launch
    traffic, traffic_src, kernel_name = None, None, "render kernel (auto-selected)"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        kernel_name = tj.get("kernel", kernel_name)
        traffic = tj["dram_bytes_per_sample"] * cn["samples"]
        traffic_src = tj.get("source")
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "fp64_pipe", "achieved": achieved_tops, "peak": peak_ginstr / 1e3, "unit": "TFLOP/s",
                "frac": achieved_tops / (peak_ginstr / 1e3) if peak_ginstr else None, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes": cn["samples"] * 32 + cn["matte"] * 24 + cn["glossy"] * 24,
                "kernel": kernel_name, "kernel_ms": last_kernel_ms,
                "ops_per_sample": ops_launch / max(1, cn["samples"]),
                "peak_source": "measured live: flux_measure_fp64_peak (unfused DADD/DMUL issue rate, FMA forbidden by parity)",
                "hbm_algorithmic_GBps": (cn["samples"] * 32 + cn["matte"] * 24) / (last_kernel_ms * 1e-3) / 1e9}

    # the memory side, against the driver-measured HBM figure (not the binding roofline: DESIGN.md §4)
    hbm_peak, hbm_src = 7700.0, "fallback: B200_PROFILING.md nominal HBM3e"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
    except (OSError, KeyError, ValueError):
        pass
    alg_gbs = roofline["algorithmic_bytes"] / (last_kernel_ms * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src,
                    "note": "not the bound: the path is FP64-issue bound (no dense contraction, 919 FP64 ops per 66-byte sample)"}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported on rank 0 at N = 1 only
            croot = args.cpu_root or pick_cpu_root(15.0)
            dt, msps = cpu_oracle_run(croot)
            cpu = {"value": msps, "unit": "Msamples/s", "cores": host_threads(), "kind": "port",
                   "sample": f"full 800x600 demo2 frame at sample_root {croot} ({croot * croot} spp), {dt:.1f} s; "
                             "C++/OpenMP oracle (Rust reference cannot be built here); Msamples/s is spp-independent",
                   "readme_reference": "5.314 Msamples/s on 44 cores (README.md:1)"}
        line = {
            "metric": "demo2.yml render throughput", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": value / 5.314, "dtype": "f64", "data": "synthetic",
            "config": workload_config(root, world), "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk, "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
            "frame_sha256": frame_sha256,
            "render_time_s_16384spp": W * H * 16384 / (value * 1e6),
            "vs_baseline_note": "value / 5.314 Msamples/s = README.md:1 (1479.9 s, 44 cores, unknown CPU)",
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
