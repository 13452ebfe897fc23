#!/usr/bin/env python
"""bench.py — headline benchmark: scenes/demo2.yml at 16384 spp (800x600, depth 5), and beside it the other
BASELINE.json configs (C1 demo1 512x512 @16 spp, C3 1M-triangle mesh @1024 spp, C4 glossy 1920x1080 @4096 spp,
C5 100 M rays x 10 K spheres) under the line's "configs" key.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--root R] [--impl reference] [--configs c1,c3,c4,c5|none]

A step is one render of the whole frame.  Our arm renders on N GPUs (one process per GPU under torchrun;
interleaved row tiles) through the C-ABI of include/fluxb200.h; every GPU's render kernel stores its pixels
straight into ONE framebuffer on rank 0's GPU over NVLink peer memory (flux_frame_*), so there is no gather
collective — a 4-byte all-reduce serves as the stream-ordered barrier.  Rank 0 prints ONE JSON line:

  value     Msamples/s, whole job, region (a) "render only": scene and sample sets resident in HBM, CUDA events,
            max over ranks, render kernels + frame assembly
  e2e       same metric, region (b) "render + job setup" through the host-buffer C-ABI: every step uploads the
            scene, generates the sample sets on the device (the reference's 1479.9 s also covers Scene::from_data
            and Camera::new, manager.rs:145-170) and reads the framebuffer back into host memory
  roofline  FP64-pipe roofline of the render kernel: algorithmic FP64 ops (SURVEY.md §8d model x device event
            counters) / kernel time, against the unfused FP64 issue rate measured live on the same GPU
            (MEASURED_PEAKS.json has no FP64 figure); ncu counters of the committed capture beside it
  cpu_baseline  the CPU oracle (a C++/OpenMP restatement of the reference; the Rust reference cannot be built in
            this image) on all host cores, on a bounded sample, both regions
  configs   C1, C3, C4, C5: value through the host-buffer C-ABI (copies inside the timed region), kernel-only
            value, roofline and a CPU baseline on a stated subset, each

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
TILE_ROWS = 1       # interleaved single rows: neighbouring rows cost the same, so the shards do (tools/tile_balance.py)
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


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_run(root: int, threads: int = 0):
    """One full-frame render of demo2 on the CPU oracle at sample_root `root`.  Returns (render seconds,
    sample-set generation seconds): region (a) is the first, region (b) — what the reference's timer covers,
    manager.rs:145-170 — their sum."""
    threads = threads or host_threads()
    from flux_b200 import JobConfiguration, SceneData
    from oracle import oracle_py as O
    sd = SceneData.from_yaml(SCENE)
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, MAX_DEPTH, 50)
    flat = sd.flatten()
    t0 = time.perf_counter()
    ss = O.generate_samples(SEED, root, MAX_DEPTH, W)          # MasterSampleSets::new inside Camera::new (single thread,
    ss.set_index = O.generate_set_index(SEED, H, W, W)         # as sampling.rs:13-33 is)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.render_rows(flat, cfg, ss, 0, H - 1, threads=threads)
    return time.perf_counter() - t0, t_gen


def pick_cpu_root(target_s: float) -> int:
    """Choose sample_root so that one oracle frame takes about target_s (per-sample cost is spp-independent)."""
    dt, _ = cpu_oracle_run(4)
    per_root1 = dt / 16.0
    return int(max(4, min(128, (target_s / per_root1) ** 0.5)))


def workload_config(gpus: int) -> dict:
    """The workload, named identically by both arms (the driver compares the two `config` objects); what an arm actually
    rendered per step is its line's `ran` object and, for the CPU arm, `cpu_baseline.sample`."""
    root = 128
    return {"workload": f"scenes/demo2.yml 800x600 at 16384 spp (sample_root 128), max_trace_depth {MAX_DEPTH}, 13 shapes (12 spheres + "
                        "1 plane), thin-lens camera; metric Msamples/s = W*H*spp / seconds",
            "sample_bound": "our arm renders the whole 16384-spp frame every step; the reference (CPU) arm renders the whole frame at a "
                            "bounded sample count per step — per-sample cost does not depend on spp — and both lines say what they "
                            "ran under `ran`",
            "sample_sets": "800 sets x (pixel CMJ, disc CMJ, 5 hemisphere MJ), seed 1 (ours: generated on the device; CPU arm: by the oracle)",
            "sharding": f"ours: interleaved tiles of {TILE_ROWS} row(s) over {gpus} GPU(s); every GPU stores its pixels into one frame on "
                        "GPU 0 over NVLink peer memory (no gather collective)",
            "l2": f"sample sets {800 * root * root * (32 + 24 * MAX_DEPTH) / 1e6:.0f} MB per GPU exceed the 126 MB L2; no cross-step "
                  "reuse of outputs"}


def ran(root: int, what: str) -> dict:
    return {"sample_root": root, "spp": root * root, "samples_per_step": 800 * 600 * root * root, "what": what}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_threads()
    root = args.cpu_root or pick_cpu_root(8.0)
    for _ in range(args.warmup):
        cpu_oracle_run(min(root, 8))
    times, gens = [], []
    for _ in range(args.steps):
        dt, tg = cpu_oracle_run(root)
        times.append(dt); gens.append(tg)
    per_step, per_gen = sum(times) / len(times), sum(gens) / len(gens)
    n = 800 * 600 * root * root
    msps, msps_b = n / per_step / 1e6, n / (per_step + per_gen) / 1e6
    sample = (f"each step = full 800x600 demo2 frame at sample_root {root} ({root * root} spp), depth {MAX_DEPTH}; "
              f"Msamples/s is spp-independent (a 16384-spp frame would take {800 * 600 * 16384 / (msps * 1e6):.0f} s)")
    line = {
        "impl": "reference", "metric": "demo2.yml render throughput", "value": msps, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": msps / 5.314, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "ran": ran(root, "CPU oracle, full 800x600 frame per step at this sample count (a bounded sample of the 16384-spp workload)"),
        "cpu_baseline": {"value": msps, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample,
                         "region": "render only (a); sample sets generated before the clock starts",
                         "value_render_plus_sample_generation": msps_b,
                         "sample_generation_s": per_gen},
        "e2e": {"value": msps, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configs
# ---------------------------------------------------------------------------------------------------------------
def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
    except (OSError, KeyError, ValueError):
        return 7700.0, "fallback: B200_PROFILING.md nominal HBM3e"


def _ncu_note(key: str):
    """Counters of the committed ncu capture of a kernel (profiles/kernel_counters.json), quoted beside the live numbers."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


class Env:
    """What the config legs share: the context, the process group, timing helpers."""

    def __init__(self, ctx, torch, dist, world, rank, dev, peak_ginstr, sm_mhz=None):
        self.ctx, self.torch, self.dist, self.world, self.rank, self.dev = ctx, torch, dist, world, rank, dev
        self.peak_ginstr = peak_ginstr
        props = torch.cuda.get_device_properties(dev)
        # L1 data-pipe peak: one wavefront per SM and cycle on each of the two paths (LSU, texture unit); the SM clock is
        # the one sampled under load during the headline (rank 0), else the device's maximum
        self.sm_count = props.multi_processor_count
        self.sm_ghz = (sm_mhz or getattr(props, "clock_rate", 1965000) / 1e3) / 1e3
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]


def bench_render_config(env: Env, name: str, sd, root: int, seed: int, steps: int, bvh: bool, cpu_fn, counter_stride: int):
    """One render config: region (a) resident (CUDA events, max over ranks), region (b) through host buffers
    (flatten + flux_set_scene [BVH build] + flux_generate_samples + render + frame to host), roofline, CPU baseline."""
    from flux_b200 import JobConfiguration
    from flux_b200.opsmodel import algorithmic_ops
    from flux_b200.sharding import FramePlan, PeerFrame
    torch, ctx = env.torch, env.ctx
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, MAX_DEPTH, 50)
    flat = sd.flatten()
    ctx.set_scene(flat, cfg)
    ctx.generate_samples(seed, W)
    plan = FramePlan(H, W, TILE_ROWS, env.world)
    pf, why = PeerFrame.try_create(ctx, plan, env.rank, env.dev, env.dist)
    fg = None
    if pf is None:   # peer-memory frame unavailable: the NCCL gather of packed slices (same fallback as the headline)
        from flux_b200.sharding import FrameGather
        fg = FrameGather(plan, env.rank, env.dev, env.dist)
    n_samples = W * H * root * root
    my_rows = plan.my_rows(env.rank)

    def step():
        if pf is not None:
            pf.render(env.stream)
            pf.barrier()
        else:
            ctx.render_row_list_device(my_rows, fg.mine.data_ptr(), env.stream)
            fg.gather()

    def read(host):
        if pf is not None:
            pf.read(host)
        else:
            host[:] = fg.frame.cpu().numpy()

    step()                                   # warm-up (these kernels are warm from the headline or trivially short)
    env.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    env.barrier()
    ms, kernel_ms = env.max_over_ranks(ev0.elapsed_time(ev1) / steps, ctx.last_kernel_ms())
    # region (b), host buffers: one step
    host_pinned = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()   # as the headline's region (b): the frame lands in pinned memory
    host = host_pinned.numpy()
    env.barrier()
    t0 = time.perf_counter()
    flat_b = sd.flatten()                    # Scene::from_data's flattening belongs to the region
    ctx.set_scene(flat_b, cfg)
    ctx.generate_samples(seed, W)
    step()
    torch.cuda.synchronize()
    if env.rank == 0:
        read(host)
    env.barrier()
    (e2e_s,) = env.max_over_ranks(time.perf_counter() - t0)
    # event counters on every counter_stride-th row of this rank's shard (the instrumented instantiation is slower)
    rows = plan.my_rows(env.rank)[::counter_stride]
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.render_row_list(rows, W)
    cn = ctx.counters()
    ctx.enable_counters(False)
    if pf is not None:
        pf.close()
    keys = sorted(cn)
    tot = dict(zip(keys, env.sum_over_ranks(*[cn[k] for k in keys])))
    scale = n_samples / max(1.0, tot["samples"])
    out = {"workload": name, "image": [W, H], "spp": root * root, "shapes": int(flat.n_shapes), "n_gpus": env.world,
           "value": n_samples / e2e_s / 1e6, "unit": "Msamples/s",
           "value_region": "host-buffer C-ABI: flatten + flux_set_scene" + (" (BVH build)" if bvh else "") +
                           " + flux_generate_samples + render + frame to pinned host memory",
           "value_resident": n_samples / (ms * 1e-3) / 1e6, "ms_per_step": ms, "kernel_ms": kernel_ms,
           "h2d_bytes_per_step": int(flat.n_shapes) * 112, "d2h_bytes_per_step": H * W * 24,
           "segments_per_sample": tot["segments"] / max(1.0, tot["samples"]),
           "frame_assembly": "peer" if pf is not None else f"nccl (peer-memory frame unavailable: {why})"}
    if not bvh:
        ops = algorithmic_ops(tot) * scale
        ach = ops / (kernel_ms * 1e-3) / 1e12 / env.world
        out["roofline"] = {"bound": "fp64_pipe", "achieved": ach, "peak": env.peak_ginstr / 1e3, "unit": "TFLOP/s",
                           "frac": ach / (env.peak_ginstr / 1e3), "ops_per_sample": ops / n_samples, "traffic": None,
                           "note": "per GPU; algorithmic FP64 ops (SURVEY.md §8d model x event counters of every "
                                   f"{counter_stride}th row) / max-over-ranks kernel time vs the live FP64 issue rate",
                           "ncu": _ncu_note("render_kernel_c1") if root * root < 64 else _ncu_note("render_wave2")}
    else:
        seg = max(1.0, tot["segments"])
        bps = (tot["nodes_visited"] * 128 + tot["bbox_tests"] * 112 + tot["tri_tests"] * 96) / seg
        hbm, src = _hbm_peak()
        gbs = bps * tot["segments"] * scale / (kernel_ms * 1e-3) / 1e9 / env.world
        out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                           "alg_bytes_per_segment": bps, "nodes_per_segment": tot["nodes_visited"] / seg,
                           "prim_tests_per_segment": (tot["bbox_tests"] + tot["tri_tests"]) / seg, "peak_source": src,
                           "note": "per GPU; algorithmic bytes = nodes x 128 + sphere records x 112 + triangle records x 96 (SURVEY.md "
                                   "§8d); served mostly from L2 — the kernel is latency / divergence bound, see ncu",
                           "ncu": _ncu_note("render_regen_bvh")}
    if env.rank == 0 and env.world == 1 and cpu_fn is not None:
        out["cpu_baseline"] = cpu_fn()
    return out


def cpu_render_subset(sd, root: int, rows, what: str, seed: int = 1):
    """The oracle on a row subset of a config's frame; Msamples/s scaled by nothing — it is a rate."""
    from flux_b200 import JobConfiguration
    from oracle import oracle_py as O
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, MAX_DEPTH, 50)
    flat = sd.flatten()
    t0 = time.perf_counter()
    ss = O.generate_samples(seed, root, MAX_DEPTH, W)
    ss.set_index = O.generate_set_index(seed, H, W, W)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.render_row_list(flat, cfg, ss, rows, threads=host_threads())
    dt = time.perf_counter() - t0
    n = len(rows) * W * root * root
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": host_threads(), "kind": "port",
            "sample": f"{what}: {len(rows)} rows x {W} px at sample_root {root} = {n} samples in {dt:.2f} s (render only; "
                      f"sample-set generation {t_gen:.2f} s on one thread)",
            "value_render_plus_sample_generation": n / (dt + t_gen) / 1e6}


def bench_c5(env: Env, n_total: int, chunk: int):
    """Config 5: n_total random rays against 10 K spheres, sharded over the ranks in contiguous chunks.  `value` goes
    through flux_trace_rays with pinned HOST buffers (ray upload and result download inside the timed region);
    `value_resident` is the kernel alone (CUDA events inside the library)."""
    from flux_b200 import JobConfiguration, synth
    torch, ctx = env.torch, env.ctx
    sd = synth.sphere_cloud_scene(10_000, seed=5)
    flat = sd.flatten()
    ctx.set_scene(flat, JobConfiguration(1))
    n_chunks = n_total // chunk
    mine = [k for k in range(n_chunks) if k % env.world == env.rank]
    o_pin = torch.empty((chunk, 3), dtype=torch.float64).pin_memory()
    d_pin = torch.empty((chunk, 3), dtype=torch.float64).pin_memory()
    hit_pin = torch.empty(chunk, dtype=torch.int32).pin_memory()
    t_pin = torch.empty(chunk, dtype=torch.float64).pin_memory()
    o_np, d_np, hit_np, t_np = o_pin.numpy(), d_pin.numpy(), hit_pin.numpy(), t_pin.numpy()
    o_dev = torch.empty((chunk, 3), dtype=torch.float64, device=env.dev)
    d_dev = torch.empty((chunk, 3), dtype=torch.float64, device=env.dev)
    hit_dev = torch.empty(chunk, dtype=torch.int32, device=env.dev)
    t_dev = torch.empty(chunk, dtype=torch.float64, device=env.dev)
    e2e_s, kernel_ms, hits, csum = 0.0, 0.0, 0, 0
    lib, C = ctx._lib, __import__("ctypes")
    from flux_b200 import _capi
    first = None
    for k in mine:
        o, d = synth.random_rays(chunk, seed=5, chunk_offset=k)      # generation is outside the timed region
        o_np[:] = o; d_np[:] = d
        if first is None:
            first = (o[:200_000].copy(), d[:200_000].copy())
        # the GPU idles (and clocks down) while the host makes the chunk: one untimed call, then the timed one
        ctx._ck(lib.flux_trace_rays(ctx._ctx, chunk, _capi.as_dp(o_np), _capi.as_dp(d_np), _capi.as_i32p(hit_np), _capi.as_dp(t_np)))
        t0 = time.perf_counter()
        ctx._ck(lib.flux_trace_rays(ctx._ctx, chunk, _capi.as_dp(o_np), _capi.as_dp(d_np), _capi.as_i32p(hit_np), _capi.as_dp(t_np)))
        e2e_s += time.perf_counter() - t0
        # the kernel alone, on the same rays resident in HBM (no copy traffic beside it): best of 2 launches
        o_dev.copy_(o_pin); d_dev.copy_(d_pin)
        torch.cuda.synchronize()
        best = None
        for _ in range(2):
            ctx.trace_rays_device(chunk, o_dev.data_ptr(), d_dev.data_ptr(), hit_dev.data_ptr(), t_dev.data_ptr(), env.stream)
            torch.cuda.synchronize()
            ms = ctx.last_kernel_ms()
            best = ms if best is None else min(best, ms)
        kernel_ms += best
        hits += int((hit_np >= 0).sum())
        csum = (csum + int(hit_np.astype(np.int64).sum())) & 0xFFFFFFFFFFFF
    e2e_max, k_max = env.max_over_ranks(e2e_s, kernel_ms)
    hits_all, = env.sum_over_ranks(hits)
    # event counts of the first chunk (instrumented instantiation, not timed)
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.trace_rays(o_np[:2_000_000], d_np[:2_000_000])
    cn = ctx.counters()
    ctx.enable_counters(False)
    seg = max(1, cn["segments"])
    bpr = (cn["nodes_visited"] * 128 + cn["bbox_tests"] * 112) / seg + 56 + 12   # + the ray itself and its result
    hbm, src = _hbm_peak()
    rays = chunk * n_chunks
    gbs = bpr * (rays / env.world) / (k_max * 1e-3) / 1e9
    # what bounds the kernel: the L1 data pipes.  An incoherent 16-byte fetch costs a pipe one wavefront
    # per lane; a node is 7 such fetches, a sphere record 7, the ray 3 and its result 1
    fpr = (cn["nodes_visited"] * 7 + cn["bbox_tests"] * 7) / seg + 4
    l1_peak = env.sm_count * env.sm_ghz * 2.0            # G wavefronts/s: LSU + texture unit
    l1_ach = fpr * (rays / env.world) / (k_max * 1e-3) / 1e9
    out = {"workload": f"{rays} random rays x 10000 spheres (BVH), ids and t per ray", "n_gpus": env.world,
           "value": rays / e2e_max / 1e6, "unit": "Mrays/s",
           "value_region": "flux_trace_rays with pinned host buffers: ray H2D + kernel + result D2H, pipelined in 4 Mi-ray pieces over three streams",
           "value_resident": rays / (k_max * 1e-3) / 1e6, "kernel_ms": k_max,
           "h2d_bytes_per_step": rays * 48, "d2h_bytes_per_step": rays * 12,
           "hit_fraction": hits_all / rays, "hit_id_checksum_rank0": csum,
           "roofline": {"bound": "l1_data_pipes", "achieved": l1_ach, "peak": l1_peak, "unit": "Gwavefronts/s", "frac": l1_ach / l1_peak,
                        "traffic": None, "fetches_per_ray": fpr, "nodes_per_ray": cn["nodes_visited"] / seg,
                        "sphere_records_per_ray": cn["bbox_tests"] / seg, "quadratics_per_ray": cn["bbox_pass"] / seg,
                        "peak_source": f"{env.sm_count} SMs x {env.sm_ghz:.3f} GHz x 2 data paths (LSU + texture unit), one wavefront "
                                       "per path and cycle (ncu l1tex__data_pipe_*_wavefronts peak)",
                        "note": "per GPU; algorithmic 16-byte fetches per ray (7 per node, 7 per sphere record, 3 + 1 for the ray and "
                                "its result) / kernel time; the tree (1.5 MB) lives in L1 / L2, so HBM is not the bound — "
                                "hbm_equivalent is the same work in bytes against the HBM figure, for SURVEY.md §8d's per-unit bytes",
                        "hbm_equivalent": {"achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "alg_bytes_per_ray": bpr,
                                           "peak_source": src},
                        "ncu": _ncu_note("trace_rays_bvh")}}
    if env.rank == 0 and env.world == 1:
        from oracle import oracle_py as O
        t0 = time.perf_counter()
        O.trace_rays(flat, first[0], first[1], threads=host_threads())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 200_000 / dt / 1e6, "unit": "Mrays/s", "cores": host_threads(), "kind": "port",
                               "sample": f"the first 200000 of the rays against all 10000 spheres by linear scan (the reference has no "
                                         f"acceleration structure, scene.rs:156-160): {dt:.2f} s"}
    return out


def run_configs(env: Env, which):
    from flux_b200 import SceneData, synth
    res = {}
    if "c1" in which:
        sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml")).with_size(512, 512)
        res["c1"] = bench_render_config(env, "scenes/demo1.yml at 512x512, 16 spp (sample_root 4), depth 5", sd, 4, 1, 5, False,
                                        lambda: cpu_render_subset(sd, 4, np.arange(512), "the whole frame"), 4)
    if "c4" in which:
        sd = synth.glossy_scene()
        res["c4"] = bench_render_config(env, "synthetic area-light + glossy/reflective scene, 67 spheres + plane, 1920x1080, 4096 spp "
                                        "(sample_root 64), depth 5", sd, 64, 4, 2, False,
                                        lambda: cpu_render_subset(sd, 6, np.arange(0, 1080, 3), "every 3rd row at reduced spp", 4), 16)
    if "c3" in which:
        sd = synth.mesh_scene(1000, 500, seed=3)
        small = synth.mesh_scene(1000, 500, seed=3, width=16, height=max(8, host_threads()))
        res["c3"] = bench_render_config(env, "synthetic 1,000,000-triangle height-field mesh (BVH), 800x600, 1024 spp (sample_root 32), depth 5",
                                        sd, 32, 3, 2, True,
                                        lambda: cpu_render_subset(small, 2, np.arange(small.output_settings.image_height),
                                                                  "linear scan over the 1 M triangles (the reference has no acceleration "
                                                                  "structure), same camera at 16 px per row", 3), 8)
    if "c5" in which:
        res["c5"] = bench_c5(env, 100_000_000, 6_250_000)   # 16 chunks: 1, 2, 4 and 8 ranks all get the same number
    return res


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
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="frame assembly: peer = kernels store into one frame over NVLink (product); nccl = all_gather of packed slices (A/B)")
    ap.add_argument("--tile-rows", type=int, default=TILE_ROWS)
    ap.add_argument("--configs", default="c1,c3,c4,c5", help="other BASELINE.json configs to run after the headline (or 'none')")
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
    from flux_b200.sharding import FrameGather, FramePlan, PeerFrame
    plan = FramePlan(H, W, args.tile_rows, world)
    my_rows = plan.my_rows(rank)
    stream = torch.cuda.current_stream().cuda_stream
    host_frame = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()
    host_np = host_frame.numpy()
    pf, assembly_note = None, None
    if args.gather == "peer":
        # the product path; if peer access / CUDA IPC is not available on this box every rank agrees to fall back to the
        # NCCL gather, and the line says so (frame_assembly)
        pf, why = PeerFrame.try_create(ctx, plan, rank, dev, dist)
        if pf is None:
            assembly_note = f"nccl (peer-memory frame unavailable: {why})"
    if pf is not None:

        def step_resident():
            pf.render(stream)
            pf.barrier()

        def frame_to_host():
            if rank == 0:
                pf.read(host_np)
    else:
        fg = FrameGather(plan, rank, dev, dist)   # packed slice (padded), gathered slices, assembled frame

        def step_resident():
            ctx.render_row_list_device(my_rows, fg.mine.data_ptr(), stream)
            fg.gather()

        def frame_to_host():
            if rank == 0:
                host_frame.copy_(fg.frame, non_blocking=False)

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
    t = torch.tensor([total_ms, last_kernel_ms, -last_kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max, kernel_ms_min = float(t[0]), float(t[1]), -float(t[2])
    ms_per_step = total_ms / args.steps
    samples_per_step = W * H * root * root
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e, region (b): host buffers through the C-ABI; scene upload + sample generation + render + D2H per step
    e2e_steps = max(1, args.e2e_steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.set_scene(flat, cfg)
        ctx.generate_samples(SEED, W)
        step_resident()
        torch.cuda.synchronize()
        frame_to_host()
        barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # the frame of the last e2e step, as rank 0 holds it: the same bytes whatever the number of GPUs (SURVEY.md §4:
    # sharding must not change per-pixel arithmetic), so the N = 1, 2, 4, 8 lines can be compared
    frame_sha256 = None
    if rank == 0:
        import hashlib
        frame_sha256 = hashlib.sha256(np.ascontiguousarray(host_np).tobytes()).hexdigest()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    scene_bytes = int(flat.struct.n_spheres * (12 * 8 + 8) + flat.struct.n_planes * (6 * 8 + 8) + flat.struct.n_materials * 56
                      + 4 * len(my_rows))
    e2e = {"value": samples_per_step / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes,
           "d2h_bytes_per_step": int(H * W * 3 * 8), "seconds_per_step": e2e_s,
           "includes": "region (b): flux_set_scene + flux_generate_samples (device) + render + framebuffer to pinned host memory"}

    # ---- roofline: algorithmic FP64 ops of one launch / kernel time vs measured FP64 issue rate
    if pf is not None:
        pf.close()
    out_dev = torch.empty((len(my_rows), W, 3), dtype=torch.float64, device=dev)
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.render_row_list_device(my_rows, out_dev.data_ptr(), stream)
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
                "hbm_algorithmic_GBps": (cn["samples"] * 32 + cn["matte"] * 24) / (last_kernel_ms * 1e-3) / 1e9,
                "ncu": _ncu_note("render_wave2")}

    # the memory side, against the driver-measured HBM figure (not the binding roofline: DESIGN.md §4)
    hbm_peak, hbm_src = _hbm_peak()
    alg_gbs = roofline["algorithmic_bytes"] / (last_kernel_ms * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src,
                    "note": "not the bound: the path is FP64-issue bound (no dense contraction, 919 FP64 ops per 66-byte sample)"}

    # ---- the other configs (every rank takes part; rank 0 reports)
    which = set() if args.configs in ("none", "") else set(args.configs.split(","))
    configs = None
    if which:
        env = Env(ctx, torch, dist, world, rank, dev, peak_ginstr, (clk or {}).get("sm_mhz"))
        ctx.set_kernel_mode(0)
        t0 = time.perf_counter()
        configs = run_configs(env, which)
        configs["seconds_spent"] = time.perf_counter() - t0

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported on rank 0 at N = 1 only
            croot = args.cpu_root or pick_cpu_root(15.0)
            dt, tg = cpu_oracle_run(croot)
            n = W * H * croot * croot
            cpu = {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": host_threads(), "kind": "port",
                   "sample": f"full 800x600 demo2 frame at sample_root {croot} ({croot * croot} spp), {dt:.1f} s; "
                             "C++/OpenMP oracle (Rust reference cannot be built here); Msamples/s is spp-independent",
                   "region": "render only (a); compare with `value`",
                   "value_render_plus_sample_generation": n / (dt + tg) / 1e6,
                   "region_b": f"render + sample-set generation ({tg:.2f} s, one thread, as MasterSampleSets::new is); compare with `e2e`",
                   "readme_reference": "5.314 Msamples/s on 44 cores (README.md:1)"}
        line = {
            "metric": "demo2.yml render throughput", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": value / 5.314, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "ran": ran(root, "the whole frame on the GPU(s), every step"), "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk, "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
            "frame_sha256": frame_sha256, "frame_assembly": assembly_note or args.gather,
            "kernel_ms_max_over_ranks": kernel_ms_max, "kernel_ms_min_over_ranks": kernel_ms_min,
            "assembly_ms_per_step": ms_per_step - kernel_ms_max,
            "render_time_s_16384spp": W * H * 16384 / (value * 1e6),
            "vs_baseline_note": "value / 5.314 Msamples/s = README.md:1 (1479.9 s, 44 cores, unknown CPU)",
            "configs": configs,
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
