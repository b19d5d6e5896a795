"""Sweep of the warp scheduler's batch parameters (track_steps x track_min_lanes).
usage: python tools/sweep_track.py [res] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudavolumerenderer_b200 import createLauncher, scenes
from cudavolumerenderer_b200.launcher import ProceduralScene

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
CASES = [("bucky", lambda: scenes.bucky(), "naiveSK"), ("hetvol", lambda: scenes.hetvol(), "regenerationSK"),
         ("manix", lambda: scenes.manix(), "regenerationSK"), ("fbm512", lambda: ProceduralScene("fbm", 512), "regenerationSK")]
for name, make, kernel in CASES:
    sc = make()
    for steps in (8, 16, 32):
        row = []
        for lanes in (8, 12, 16, 20, 24, 28):
            kl = createLauncher(kernel, 0, track_steps=steps, track_min_lanes=lanes)
            kl.setScene(sc)
            best = None
            for rep in range(2):
                kl.resetCounters()
                kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
                c = kl.counters()
                if best is None or c["kernel_ms"] < best["kernel_ms"]:
                    best = c
            row.append(f"{res * res * spp / best['kernel_ms'] / 1e3:7.0f}")
            kl.close()
        print(f"{name:8s} steps {steps:3d} | min_lanes 8/12/16/20/24/28: " + " ".join(row), flush=True)
