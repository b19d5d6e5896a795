"""Throughput of the five BASELINE.json configurations on one GPU (device-timed by the
library's CUDA events around every launch; the default bench.py line is config 2).
usage: python tools/bench_configs.py [spp_scale]   (spp_scale < 1 shortens the run)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import cudavolumerenderer_b200 as cvr  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
CONFIGS = [
    ("C1 bucky 32^3, 256x256, 16 spp, 1 tile, naiveSK", "naiveSK", lambda: cvr.scenes.bucky(), 256, 16, (1, 1), {}),
    ("C2 hetvol 128x128x50, 1024x1024, 64 spp, regenerationSK", "regenerationSK", lambda: cvr.scenes.hetvol(), 1024, 64, (1, 1), {}),
    ("C3 manix 256x230x256, 1024x1024, 256 spp, 10x10 tiles, regenerationSK", "regenerationSK", lambda: cvr.scenes.manix(), 1024, 256, (10, 10), {}),
    ("C4 fbm 1024^3 albedo 0.99, 2048x2048, 128 spp", "regenerationSK", lambda: cvr.scenes.fbm_device(1024), 2048, 128, (1, 1), {}),
    ("C5 sparse 2048^3, 4096x4096, 1024 spp (run at 16 spp), global majorant", "regenerationSK", lambda: cvr.scenes.sparse_fbm(2048), 4096, 16, (8, 8), {}),
    ("C5 sparse 2048^3, 4096x4096, 16 spp, tracking=local", "regenerationSK", lambda: cvr.scenes.sparse_fbm(2048), 4096, 16, (8, 8), {"tracking": "local"}),
]
out = []
for name, kernel, mk, res, spp, tiles, opts in CONFIGS:
    spp = max(1, int(round(spp * scale)))
    sc = mk()
    kl = cvr.createLauncher(kernel, 0, **opts)
    kl.setScene(sc)
    for rep in range(2):  # first repetition warms up
        kl.resetCounters()
        kl.setSeed(0)
        img = kl.renderImage((res, res), tiles, spp, fov_x=sc.fov_x, fuse_tiles=True)
        c = kl.counters()
    alg = 32 * c["density_lookups"] + 128 * (0 if getattr(sc, "albedo", None) is None else c["albedo_lookups"]) + 16 * c["paths"]
    info = kl.volumeInfo()
    row = {"config": name, "spp_run": spp, "paths": c["paths"], "kernel_ms": c["kernel_ms"], "launches": c["launches"],
           "msamples_per_s": c["paths"] / c["kernel_ms"] / 1e3, "density_lookups_per_s": c["density_lookups"] / c["kernel_ms"] * 1e3,
           "algorithmic_gb_per_s": alg / c["kernel_ms"] / 1e6, "lookups_per_path": c["density_lookups"] / c["paths"],
           "bounces_per_path": c["bounces"] / c["paths"], "speculative_lookups": c["speculative_lookups"],
           "skipped_fetches": c["skipped_fetches"],
           "fetched_fraction": 1.0 - c["skipped_fetches"] / max(c["density_lookups"] + c["speculative_lookups"], 1),
           "layout": info["layout"], "layout_gb": info["layout_bytes"] / 1e9, "image_mean": float(np.nanmean(img[..., :3])),
           "options": {k: kl.getOption(k) for k in ("sched", "warp_slots", "pair", "tracking", "exact", "skip")},
           "launch_shape": list(kl.launchShape())}
    out.append(row)
    print(json.dumps(row), flush=True)
    kl.close()
