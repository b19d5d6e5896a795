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
roofline   dominant kernel k_volpt: algorithmic bytes per launch (32 B/density lookup +
           128 B/albedo lookup + 16 B/path; the counts come from the
           kernel's own counters) / its mean launch duration (CUDA events recorded by the
           library around every launch) against MEASURED_PEAKS.json's hbm_gbs.
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
        "config": {"workload": WORKLOAD, "sample": f"{RES}x{RES} at {spp} spp per step"},
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
    d_img = torch.zeros((RES, RES, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    from cudavolumerenderer_b200.distributed import render_sharded, spp_shard

    first, count = spp_shard(total_spp, rank, world)

    def step_device():
        flush.fill_(1)  # L2 flush between timed iterations
        kl.setSeed(0)
        # this rank's 64 sample indices of the 64*N-spp image, then ONE NCCL all-reduce
        render_sharded(kl, (RES, RES), (1, 1), total_spp, "spp", d_img, fov_x=sc.fov_x)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

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
    ms = e0.elapsed_time(e1)
    ctr = kl.counters()
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
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
        kl.setScene(sc_pinned)  # H2D of the volumes + device layout build
        kl.setSeed(0)
        if dist is None:
            kl.renderImage((RES, RES), (1, 1), total_spp, fov_x=sc.fov_x, host_image=host_img.numpy())  # D2H inside
        else:
            render_sharded(kl, (RES, RES), (1, 1), total_spp, "spp", d_img, fov_x=sc.fov_x)
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
    ms_e2e = max(e0.elapsed_time(e1), wall)  # host-side copies are part of the call: take the wall clock
    clk.stop()
    if dist is not None:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = paths_per_step * n_e2e / ms_e2e / 1e3
    h2d = int(sc.density.nbytes + sc.albedo.nbytes + 12 * 4 + 2 * 4)
    d2h = int(RES * RES * 16)

    if rank != 0:
        kl.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (this rank's launches in the timed region)
    launches = int(ctr["launches"])
    alg_bytes = 32 * ctr["density_lookups"] + 128 * ctr["albedo_lookups"] + 16 * ctr["paths"]
    kern_ms = ctr["kernel_ms"] / max(launches, 1)
    peak, peak_src = peaks()
    achieved = alg_bytes / max(launches, 1) / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("k_volpt_dram_bytes_per_launch")
    except Exception:
        pass
    grid, block, regs = kl.launchShape()
    # measured random-32-B-sector gather peak at this volume's device footprint (L2-resident here)
    nz, ny, nx = sc.density.shape
    footprint = (nx + 1) * (ny + 1) * (nz + 1) * (32 + 96)  # density cells + rgb albedo cells
    gather_peak = kl.gatherRoofline(footprint, 512, 8)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": {"warp": "k_volpt_warp", "queued": "k_volpt_queued", "sorted": "k_volpt_sorted"}.get(kl.getOption("sched"), "k_volpt"),
                "peak_source": peak_src,
                "kernel_ms_per_launch": kern_ms, "algorithmic_bytes_per_launch": alg_bytes / max(launches, 1),
                "density_lookups_per_s": ctr["density_lookups"] / (ctr["kernel_ms"] * 1e-3),
                "kernel_share_of_step": ctr["kernel_ms"] / ms,
                "launch": {"grid": grid, "block": block, "regs": regs},
                "gather_roofline": {"what": "random 32-B-sector 256-bit gather microbenchmark at the volume's device footprint",
                                    "footprint_bytes": footprint, "peak": gather_peak, "unit": "GB/s",
                                    "frac": achieved / gather_peak}}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    ref_gpu = None
    if world == 1:
        cpu_spp = 16
        v, cores, dt, _, kind, what = cpu_sample(sc, cpu_spp)
        cpu = {"value": v, "unit": METRIC, "cores": cores, "kind": kind,
               "sample": f"{RES}x{RES} at {cpu_spp} spp ({dt:.1f} s), {what}"}
        ref_gpu = ref_gpu_baseline(sc, SPP, rgb.float().cpu().numpy())

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "resolution": [RES, RES], "spp_per_gpu": SPP, "total_spp": total_spp,
                   "kernel": "regenerationSK", "rng": kl.getOption("rng"), "layout": kl.getOption("layout"),
                   "sched": kl.getOption("sched"), "arithmetic": "fused (exact=0)" if kl.getOption("exact") == "0" else "reference order (exact=1)",
                   "tracking": kl.getOption("tracking"), "warp_slots": kl.getOption("warp_slots"),
                   "speculative_pair_step": kl.getOption("pair"), "fetch_skip_table": kl.getOption("skip"),
                   "exit_others": kl.getOption("exit_others"),
                   "sharding": "spp" if world > 1 else "none", "l2": "flushed between steps (256 MiB write)",
                   "image_mean": img_mean, "nan_pixels": nan_px},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / n_e2e},
        "gpu_launches": launches,
        "roofline": roofline,
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
