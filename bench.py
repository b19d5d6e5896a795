#!/usr/bin/env python
"""bench.py -- throughput of the volumetric path-tracing hot path (Msamples/s).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one render of the workload BASELINE.json's metric is quoted on:
configs[1], "hetvol heterogeneous medium 1024x1024, 64 spp, regenerationSK on 1 B200"
(procedural stand-in for the LFS-stub payload).  With N GPUs the
job is spp-sharded (weak scaling: every rank renders its own 64 sample indices of a
64*N-spp image with the volume replicated) and the framebuffer is combined with one
NCCL all-reduce inside the timed region.

value      whole-job Msamples/s, scene resident in HBM, device time (CUDA events on the
           launching stream, max over ranks), L2 flushed between steps.
e2e        same metric through the public C-ABI call with HOST buffers: per step the
           volumes are uploaded from pinned host memory (cvr_set_scene), the image is
           rendered (cvr_render_image) and read back into host memory.
roofline   dominant kernel k_volpt_warp.  `achieved` / `frac` keep SURVEY.md 8(d)'s definition --
           algorithmic bytes per launch (32 B/density lookup + 128 B/albedo lookup + 16 B/path; the
           counts come from the kernel's own counters) / its mean launch duration (CUDA events
           recorded by the library around every launch) against MEASURED_PEAKS.json's hbm_gbs --
           and `bound` says what the launch is really limited by: "l2-gather" when the device
           layout fits the L2 (then `gather` = L2 -> SM sector traffic against the measured
           random-sector gather peak at that footprint is the utilisation figure), "hbm" beyond.
configs    the same measurement for C3 (MANIX 1024^2 x 256 spp, 10x10 tiles: north_star's target)
           and C4 (fBm 1024^3, 2048^2 x 128 spp: the HBM-resident dense volume), N = 1 only.
strong_scaling  C3 with the image and spp FIXED, sharded over the N ranks with the balanced plan
           (cvr_shard_plan: the sample split, since 256 spp divide evenly) and ONE reduce to rank 0,
           device-timed, max over ranks.
cpu_baseline  the reference's OWN regenerationSK kernel (d_render_single_thread_regeneration
           and everything it calls) compiled for the host by g++ from the reference's headers
           (oracle/_ref/libcvr_ref_cpu.so, kind "reference": one persistent CUDA thread per host
           core sharing the reference's atomic path counter) on all host cores over a bounded
           sample of the workload; the plain-C oracle (kind "port") only where that library is
           absent.  The reference has no CPU renderer of its own.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s"
RES = 1024
SPP = 64
WORKLOAD = "hetvol 128x128x50 (procedural stand-in), 1024x1024, 64 spp/GPU, regenerationSK, 1 tile"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions: ONE nvidia-smi process in
    loop mode (-lms 200), started before the timed region and killed at the end of the run, as in
    the profiling recipe.  (Spawning one nvidia-smi per sample stalled CUDA calls of this process
    for 100-500 ms whenever one of them exited: seen as outliers among the 66 ms e2e steps.)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []      # (time, fields)
        self.window = None  # (t0, t1) of the device-timed region
        self._p = None
        self._t = None

    def _run(self):
        try:
            for line in self._p.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    self.rows.append((time.perf_counter(), f))
        except Exception:
            pass

    def start(self):
        if os.environ.get("CVR_BENCH_NO_CLOCKS"):  # diagnosis only
            return self
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                        "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self._p = None
        return self

    def stop(self):
        if self._p is not None:
            try:
                self._p.terminate()
                self._p.wait(timeout=5)
            except Exception:
                pass
            self._p = None

    def __enter__(self):  # marks the device-timed region
        if self._p is None:
            self.start()
        self._w0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.window = (self._w0, time.perf_counter())

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        rows = [f for t, f in self.rows if self.window is None or self.window[0] - 0.25 <= t <= self.window[1] + 0.25]
        if not rows:
            rows = [f for _, f in self.rows]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_scene_and_cam(sc, res):
    from oracle import bindings as B

    osc = B.make_scene(sc.density, sc.albedo, sc.box_min, sc.box_max, sc.scale, sc.max_density)
    cam = B.make_camera(res, res, res, res, fov_x=sc.fov_x)
    return B, osc, cam


def cpu_sample(sc, spp: int, seed: int = 0):
    """`spp` samples of the full 1024^2 image on all host cores: the reference's own
    regenerationSK kernel built for the host when oracle/_ref/libcvr_ref_cpu.so is present
    (kind "reference"), else the plain-C oracle (kind "port").
    Returns (Msamples/s, cores, seconds, paths, kind, what)."""
    B, osc, cam = oracle_scene_and_cam(sc, RES)
    cores = os.cpu_count() or 1
    if B.ref_cpu() is not None:
        rc = B.RefCpu(osc, cam)
        t0 = time.perf_counter()
        rc.render_regen(spp, seed=seed, n_threads=cores)
        dt = time.perf_counter() - t0
        paths = RES * RES * spp
        return paths / dt / 1e6, cores, dt, paths, "reference", \
            "the reference's d_render_single_thread_regeneration compiled for the host (oracle/_ref/libcvr_ref_cpu.so)"
    t0 = time.perf_counter()
    _, ctr = B.render_regen(osc, cam, spp, seed=seed, rng_mode=1, n_threads=cores)
    dt = time.perf_counter() - t0
    return ctr["paths"] / dt / 1e6, cores, dt, ctr["paths"], "port", "oracle/cvr_oracle.c regenerationSK path loop"


def run_reference(args, rank: int):
    """Reference arm: the reference has no CPU renderer and cannot be pip-installed (it
    is a CMake/vcpkg C++ executable), so this arm times the reference's own
    regenerationSK kernel compiled for the host from its headers (oracle/_ref/libcvr_ref_cpu.so;
    the plain-C oracle where that is absent) on all host cores; each step is a bounded sample
    (1024x1024 at 2 spp) of the same workload."""
    if rank != 0:
        return
    from cudavolumerenderer_b200 import scenes

    sc = scenes.hetvol()
    spp = 2
    for _ in range(args.warmup):
        cpu_sample(sc, spp)
    t0 = time.perf_counter()
    paths = 0
    kind = what = None
    for s in range(args.steps):
        _, cores, _, n, kind, what = cpu_sample(sc, spp, seed=s * RES * RES * spp)
        paths += n
    dt = time.perf_counter() - t0
    v = paths / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{RES}x{RES} at {spp} spp per step", "spp_scaled": True,
                   "note": "a rate metric on a bounded sample: the CPU renders 2 of the 64 spp per step"},
        "cpu_baseline": {"value": v, "unit": METRIC, "cores": cores, "kind": kind,
                         "sample": f"{RES}x{RES} at {spp} spp per step, {what}"},
        "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def ref_gpu_baseline(sc, spp: int, ours_rgb=None):
    """The reference's OWN regenerationSK(thread) kernel compiled from its headers for
    sm_100a (oracle/_ref/libcvr_ref_gpu.so), same scene / resolution: context number."""
    import ctypes as C

    from oracle import bindings as B

    if not os.path.exists(B.REF_GPU_SO):
        return None
    try:
        from cudavolumerenderer_b200 import abi

        R = C.CDLL(B.REF_GPU_SO)
        R.refgpu_last_error.restype = C.c_char_p
        nz, ny, nx = sc.density.shape
        bmin, bmax = (C.c_float * 3)(*sc.box_min), (C.c_float * 3)(*sc.box_max)
        if R.refgpu_set_scene(sc.density.ctypes.data_as(C.c_void_p), nx, ny, nz,
                              sc.albedo.ctypes.data_as(C.c_void_p), nx, ny, nz, bmin, bmax,
                              C.c_float(sc.scale), C.c_float(sc.max_density)):
            return None
        iv, rtv = abi.default_camera(RES, RES, sc.fov_x)
        if R.refgpu_set_camera((C.c_float * 12)(*iv.tolist()), (C.c_float * 2)(*rtv.tolist()), RES, RES,
                               C.c_float(RES), C.c_float(RES), 0, 0):
            return None
        import numpy as np

        best = None
        ms, g, b = C.c_float(), C.c_int(), C.c_int()
        imgs = []
        for i in range(3):
            out = np.zeros((RES, RES, 4), np.float32)
            if R.refgpu_render(1, spp, 1000 + i, out.ctypes.data_as(C.c_void_p), C.byref(ms), C.byref(g), C.byref(b)):
                return None
            best = ms.value if best is None else min(best, ms.value)
            imgs.append(out[..., :3] / spp)
        R.refgpu_release()
        res = {"kernel": "reference regenerationSK(thread) recompiled for sm_100a", "ms": best,
               "value": RES * RES * spp / best / 1e3, "unit": METRIC, "grid": g.value, "block": b.value}

        def rel_rmse(a, r):  # BASELINE "RMSE vs ref": sqrt(mean((I-R)^2)) / mean(R) over RGB, NaN pixels left out
            ok = ~(np.isnan(a).any(axis=-1) | np.isnan(r).any(axis=-1))
            return float(np.sqrt(np.mean((a[ok] - r[ok]) ** 2)) / np.mean(r[ok]))

        if ours_rgb is not None:
            # images at matched spp from independent streams: agreement = the same relative RMSE as
            # two runs of the reference kernel against each other (pure Monte-Carlo noise)
            res["rmse_vs_ref"] = {"spp": spp, "ours_vs_reference": rel_rmse(ours_rgb, imgs[0]),
                                  "reference_vs_reference": rel_rmse(imgs[1], imgs[0]),
                                  "mean_ours": float(np.nanmean(ours_rgb)), "mean_reference": float(np.nanmean(imgs[0]))}
        return res
    except Exception as e:  # context number only
        return {"error": str(e)}


def traffic_profile(name: str):
    """ncu-measured hit rates / DRAM traffic of one launch of this workload (profiles/traffic.json, written by hand
    from the ncu exports named in its "source" fields): they turn the in-run counters into L2 -> SM and DRAM sector
    traffic.  None when the workload was not profiled."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("workloads", {}).get(name)
    except Exception:
        return None


def roofline_of(kl, ctr, step_ms, workload: str, const_albedo: bool, footprint_bytes: int, l2_bytes: int):
    """Roofline object of the launches counted in `ctr` (cvr_counters since the last reset).

    algorithmic bytes = 32 B x density lookups + 128 B x albedo lookups (0 for a constant albedo)
    + 16 B x paths: layout-independent, what the reference's algorithm asks for.
    requested bytes = what the kernel's load instructions ask the L1 for: 32 B x cell loads issued (lookups +
    speculative - skipped) + 96 B x albedo lookups (three 256-bit loads of an rgb cell) + 16 B x paths.
    L2 -> SM and DRAM sector traffic = requested x the miss rates ncu measured on this workload (profiles/)."""
    launches = max(int(ctr["launches"]), 1)
    kern_ms = ctr["kernel_ms"] / launches
    n_alb = 0 if const_albedo else ctr["albedo_lookups"]
    alg = (32 * ctr["density_lookups"] + 128 * n_alb + 16 * ctr["paths"]) / launches
    issued = ctr["density_lookups"] + ctr["speculative_lookups"] - ctr["skipped_fetches"]
    req = (32 * issued + 96 * n_alb + 16 * ctr["paths"]) / launches
    peak, peak_src = peaks()
    achieved = alg / (kern_ms * 1e-3) / 1e9
    l2_resident = footprint_bytes <= l2_bytes
    prof = traffic_profile(workload)
    grid, block, regs = kl.launchShape()
    r = {"bound": "l2-gather" if l2_resident else "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": prof.get("dram_bytes_per_launch") if prof else None,
         "traffic_source": prof.get("source") if prof else None,
         "what": "achieved = ALGORITHMIC bytes / kernel time (SURVEY.md 8(d)); frac = achieved / measured HBM copy peak -- a "
                 "throughput figure, not a utilisation: cache reuse among the 8 corners of a cell and L1 / L2 hits put it "
                 "above the bytes that move.  The utilisation figures are `gather` (L2-resident) and `dram` (HBM-resident).",
         "kernel": "k_volpt_warp", "peak_source": peak_src, "kernel_ms_per_launch": kern_ms,
         "algorithmic_bytes_per_launch": alg, "requested_bytes_per_launch": req,
         "requested_gbs": req / (kern_ms * 1e-3) / 1e9,
         "cell_loads_issued_fraction": issued / max(ctr["density_lookups"] + ctr["speculative_lookups"], 1),
         "density_lookups_per_s": ctr["density_lookups"] / (ctr["kernel_ms"] * 1e-3),
         "kernel_share_of_step": (ctr["kernel_ms"] / step_ms) if step_ms else None,
         "launch": {"grid": grid, "block": block, "regs": regs},
         "device_layout_bytes": footprint_bytes, "l2_bytes": l2_bytes}
    # the measured gather ceilings at this footprint: random 32-byte sectors, 8 independent 256-bit loads in flight per
    # thread, with the L1 allowed (hashed indices: no reuse) and bypassed (ld.global.nc.L1::no_allocate)
    fp = min(max(footprint_bytes, 1 << 20), 64 << 30)
    g_l1 = kl.gatherRoofline(fp, 512, 8)
    g_nol1 = kl.gatherRoofline(fp, 512, 8, bypass_l1=True)
    gpeak = max(g_l1, g_nol1)
    gather = {"what": "random 32-B-sector 256-bit gather microbenchmark at the device layout's footprint (GB/s)",
              "peak_l1_allowed": g_l1, "peak_l1_bypassed": g_nol1, "peak": gpeak, "unit": "GB/s"}
    if prof and prof.get("l1_sector_hit_rate") is not None:
        l2_to_sm = req * (1.0 - prof["l1_sector_hit_rate"])
        gather.update({"l2_to_sm_gbs": l2_to_sm / (kern_ms * 1e-3) / 1e9, "l1_sector_hit_rate": prof["l1_sector_hit_rate"],
                       "frac": l2_to_sm / (kern_ms * 1e-3) / 1e9 / gpeak,
                       "frac_what": "L2 -> SM sector traffic (requested bytes x ncu's L1 sector miss rate) / gather peak"})
    else:
        gather.update({"frac": None, "frac_what": "no ncu profile of this workload under profiles/: requested_gbs / peak would count L1 hits"})
    r["gather"] = gather
    if prof and prof.get("dram_bytes_per_launch_at") and not l2_resident:
        # DRAM sector traffic scales with the lookups of the same scene: bytes per density lookup from the profiled launch
        per_lookup = prof["dram_bytes_per_launch_at"]["dram_bytes"] / prof["dram_bytes_per_launch_at"]["density_lookups"]
        dram = per_lookup * ctr["density_lookups"] / launches
        r["traffic"] = dram  # per launch of THIS run, like `achieved`
        r["traffic_source"] = prof.get("source", "") + " -- DRAM bytes per density lookup of that capture x the lookups counted in this run"
        r["dram"] = {"dram_gbs": dram / (kern_ms * 1e-3) / 1e9, "dram_bytes_per_density_lookup": per_lookup,
                     "frac_of_hbm_copy_peak": dram / (kern_ms * 1e-3) / 1e9 / peak,
                     "what": "DRAM read+write bytes per density lookup from the ncu capture x density lookups of this run; the "
                             "binding resource of an HBM-resident gather is not bytes but L2-missing REQUESTS (44.7 G/s measured, "
                             "profiles/r2_dram_granule.md): `miss_requests`",
                     "miss_requests_per_s": issued / launches * (1.0 - prof.get("l1_sector_hit_rate", 0.0)) *
                                            (1.0 - prof.get("l2_sector_hit_rate", 0.0)) / (kern_ms * 1e-3),
                     "miss_request_ceiling_per_s": 44.7e9}
        r["dram"]["frac_of_miss_request_ceiling"] = r["dram"]["miss_requests_per_s"] / r["dram"]["miss_request_ceiling_per_s"]
    return r


def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch

    from cudavolumerenderer_b200 import RegenerationVolPTsk, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its banner ("NCCL version ...") on stdout when the first communicator comes up;
        # stdout carries the ONE JSON line, so file descriptor 1 points at stderr until NCCL is up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            dist.reduce(warm, dst=0)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    sc = scenes.hetvol()
    total_spp = SPP * world
    kl = RegenerationVolPTsk(local_rank)
    stream = torch.cuda.current_stream(dev)
    kl.setStream(stream.cuda_stream)
    kl.setScene(sc)
    l2_bytes = int(torch.cuda.get_device_properties(dev).L2_cache_size)
    d_img = torch.zeros((RES, RES, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    from cudavolumerenderer_b200.distributed import render_sharded, spp_shard

    first, count = spp_shard(total_spp, rank, world)

    def step_device():
        flush.fill_(1)  # L2 flush between timed iterations
        kl.setSeed(0)
        # this rank's 64 sample indices of the 64*N-spp image, then ONE NCCL reduce to rank 0 (the rank that
        # keeps the image; an all-reduce would move twice the bytes for nothing)
        render_sharded(kl, (RES, RES), (1, 1), total_spp, "spp", d_img, fov_x=sc.fov_x, reduce_to=0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_device()
    barrier()
    kl.resetCounters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ctr = kl.counters()
    paths_per_step = RES * RES * SPP * world
    value = paths_per_step * args.steps / ms / 1e3
    rgb = d_img[..., :3]
    nan_px = int(torch.isnan(rgb).any(dim=2).sum().item())  # the reference's own u == 1.0 GGX singularity
    img_mean = float(torch.nanmean(rgb).item())

    # ---- end-to-end through the C ABI with host buffers
    pin_den = torch.from_numpy(sc.density).pin_memory()
    pin_alb = torch.from_numpy(sc.albedo).pin_memory()
    host_img = torch.zeros((RES, RES, 4), dtype=torch.float32).pin_memory()
    sc_pinned = type(sc)(pin_den.numpy(), pin_alb.numpy(), sc.box_min, sc.box_max, sc.scale, sc.max_density,
                         fov_x=sc.fov_x, name=sc.name)

    def step_e2e():
        flush.fill_(1)
        kl.setScene(sc_pinned)  # H2D of the volumes + device layout build (every rank holds a replica)
        kl.setSeed(0)
        if dist is None:
            kl.renderImage((RES, RES), (1, 1), total_spp, fov_x=sc.fov_x, host_image=host_img.numpy())  # D2H inside
        else:
            render_sharded(kl, (RES, RES), (1, 1), total_spp, "spp", d_img, fov_x=sc.fov_x, reduce_to=0)
            if rank == 0:  # the image lands on the one rank that keeps it
                host_img.copy_(d_img, non_blocking=False)

    step_e2e()
    barrier()
    n_e2e = max(2, args.steps)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(n_e2e):
        ts = time.perf_counter()
        step_e2e()
        if os.environ.get("CVR_BENCH_DEBUG"):
            print(f"[bench] e2e step {(time.perf_counter() - ts) * 1e3:.2f} ms", file=sys.stderr)
    e1.record(stream)
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall))  # host-side copies are part of the call: take the wall clock
    e2e_value = paths_per_step * n_e2e / ms_e2e / 1e3
    h2d = int(sc.density.nbytes + sc.albedo.nbytes + 12 * 4 + 2 * 4)
    d2h = int(RES * RES * 16)

    # ---- roofline of the dominant kernel (this rank's launches in the timed region)
    nz, ny, nx = sc.density.shape
    footprint = (nx + 1) * (ny + 1) * (nz + 1) * (32 + 96)  # density cells + rgb albedo cells
    roofline = roofline_of(kl, ctr, ms, "hetvol", False, footprint, l2_bytes) if rank == 0 else None
    main_opts = {k: kl.getOption(k) for k in ("rng", "layout", "sched", "exact", "tracking", "warp_slots", "pair", "skip", "exit_others")}
    launches = int(ctr["launches"])

    # ---- strong scaling of the tiled configuration (C3: MANIX 1024^2 x 256 spp, 10 x 10 tiles), every N
    C3_RES, C3_SPP, C3_TILES = 1024, 256, (10, 10)
    msc = scenes.manix()
    kl.setScene(msc)
    d_img3 = torch.zeros((C3_RES, C3_RES, 4), dtype=torch.float32, device=dev)

    def step_c3():
        flush.fill_(1)
        kl.setSeed(0)
        render_sharded(kl, (C3_RES, C3_RES), C3_TILES, C3_SPP, "balanced", d_img3, fov_x=msc.fov_x, reduce_to=0)

    for _ in range(2):
        step_c3()
    barrier()
    kl.resetCounters()
    n_c3 = max(3, min(args.steps, 10))
    e0.record(stream)
    for _ in range(n_c3):
        step_c3()
    e1.record(stream)
    barrier()
    ms_c3 = max_over_ranks(e0.elapsed_time(e1)) / n_c3
    ctr3 = kl.counters()
    c3_paths = 102 * 102 * 100 * C3_SPP  # Q6: 10 x 10 tiles of 102^2 cover 1020^2 pixels
    strong = {"workload": "C3 MANIX-like 256x230x256 (procedural stand-in), 1024x1024, 256 spp, 10x10 tiles (1020^2 pixels covered), regenerationSK",
              "scaling": "strong", "n_gpus": world, "sharding": "balanced (cvr_shard_plan: 256 samples over N ranks = the sample split, every rank traces the same number of paths through every tile in one launch), one reduce to rank 0",
              "ms_per_step": ms_c3, "value": c3_paths / ms_c3 / 1e3, "unit": METRIC, "steps": n_c3,
              "image_mean": float(torch.nanmean(d_img3[:1020, :1020, :3]).item()) if rank == 0 else None}

    if rank != 0:
        kl.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    configs = None
    cpu = None
    ref_gpu = None
    if world == 1:
        # ---- the other single-GPU configurations of BASELINE.json on the same line: C3 (above) and C4
        mnz, mny, mnx = msc.density.shape
        fp3 = (mnx + 1) * (mny + 1) * (mnz + 1) * (32 + 96)
        c3 = {"config": "C3", "workload": strong["workload"], "value": strong["value"], "unit": METRIC, "ms_per_step": ms_c3,
              "steps": n_c3, "options": {k: kl.getOption(k) for k in ("warp_slots", "skip", "pair")},
              "lookups_per_path": ctr3["density_lookups"] / max(ctr3["paths"], 1),
              "roofline": roofline_of(kl, ctr3, ms_c3 * n_c3, "manix", False, fp3, l2_bytes)}
        del d_img3
        fsc = scenes.fbm_device(1024)
        kl.setScene(fsc)
        C4_RES, C4_SPP = 2048, 128
        d_img4 = torch.zeros((C4_RES, C4_RES, 4), dtype=torch.float32, device=dev)
        for _ in range(2):
            kl.setSeed(0)
            kl.renderImage((C4_RES, C4_RES), (1, 1), C4_SPP, fov_x=fsc.fov_x, d_image=d_img4.data_ptr())
        torch.cuda.synchronize(dev)
        kl.resetCounters()
        n_c4 = 3
        e0.record(stream)
        for _ in range(n_c4):
            kl.setSeed(0)
            kl.renderImage((C4_RES, C4_RES), (1, 1), C4_SPP, fov_x=fsc.fov_x, d_image=d_img4.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms_c4 = e0.elapsed_time(e1) / n_c4
        ctr4 = kl.counters()
        c4 = {"config": "C4", "workload": "C4 fBm 1024^3 (generated on the device, 34.5 GB of lookup cells), albedo 0.99, 2048x2048, 128 spp, regenerationSK",
              "value": C4_RES * C4_RES * C4_SPP / ms_c4 / 1e3, "unit": METRIC, "ms_per_step": ms_c4, "steps": n_c4,
              "options": {k: kl.getOption(k) for k in ("warp_slots", "skip", "pair")},
              "lookups_per_path": ctr4["density_lookups"] / max(ctr4["paths"], 1),
              "bounces_per_path": ctr4["bounces"] / max(ctr4["paths"], 1),
              "image_mean": float(torch.nanmean(d_img4[..., :3]).item()),
              "roofline": roofline_of(kl, ctr4, ms_c4 * n_c4, "fbm1024", True, kl.volumeInfo()["layout_bytes"], l2_bytes)}
        del d_img4
        configs = [c3, c4]
        kl.setScene(sc)
        # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
        # ~10-15 s of CPU work: the whole 64 spp on a 32-core box (5 Msamples/s), half of them on 16 cores (2.6 Msamples/s)
        cpu_spp = SPP if (os.cpu_count() or 1) >= 32 else SPP // 2
        v, cores, dt, _, kind, what = cpu_sample(sc, cpu_spp)
        cpu = {"value": v, "unit": METRIC, "cores": cores, "kind": kind,
               "sample": f"{RES}x{RES} at {cpu_spp} spp ({dt:.1f} s), {what}"}
        ref_gpu = ref_gpu_baseline(sc, SPP, rgb.float().cpu().numpy())
    clk.stop()

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "resolution": [RES, RES], "spp_per_gpu": SPP, "total_spp": total_spp,
                   "kernel": "regenerationSK", "rng": main_opts["rng"], "layout": main_opts["layout"],
                   "sched": main_opts["sched"], "arithmetic": "fused (exact=0)" if main_opts["exact"] == "0" else "reference order (exact=1)",
                   "tracking": main_opts["tracking"], "warp_slots": main_opts["warp_slots"],
                   "speculative_pair_step": main_opts["pair"], "fetch_skip_table": main_opts["skip"],
                   "exit_others": main_opts["exit_others"],
                   "sharding": "spp, one ncclReduce of the framebuffer to rank 0" if world > 1 else "none",
                   "l2": "flushed between steps (256 MiB write)",
                   "image_mean": img_mean, "nan_pixels": nan_px},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / n_e2e},
        "gpu_launches": 2 * launches,
        "gpu_launches_what": f"rank 0, timed region: {launches} x k_volpt_warp (the path kernel, counted by the library) + "
                             f"{launches} x k_resolve_tiles (one fused resolve per render); the L2 flush fill and the image "
                             "zero fill are torch kernels and not counted",
        "roofline": roofline,
        "configs": configs,
        "strong_scaling": strong,
        "cpu_baseline": cpu,
        "reference_gpu_kernel": ref_gpu,
        "density_lookups_per_s": roofline["density_lookups_per_s"],
    }
    print(json.dumps(line), flush=True)
    kl.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # not under torchrun: re-launch ourselves on N GPUs of this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
