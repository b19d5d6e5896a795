"""A/B of the fetch-skip table (option "skip") on the usual scenes: kernel ms, Msamples/s and the
fraction of density fetches skipped.  usage: python tools/ab_skip.py [res] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudavolumerenderer_b200 import createLauncher, scenes
from cudavolumerenderer_b200.launcher import ProceduralScene

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
CASES = [("bucky", lambda: scenes.bucky(), "naiveSK"), ("hetvol", lambda: scenes.hetvol(), "regenerationSK"),
         ("manix", lambda: scenes.manix(), "regenerationSK"), ("fbm512", lambda: ProceduralScene("fbm", 512), "regenerationSK"),
         ("sparse1024", lambda: ProceduralScene("sparsefbm", 1024), "regenerationSK")]
for name, make, kernel in CASES:
    sc = make()
    for opts in ({"skip": 0}, {"skip": 1}, {}):
        kl = createLauncher(kernel, 0, **opts)
        kl.setScene(sc)
        best = None
        for rep in range(3):
            kl.resetCounters()
            kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
            c = kl.counters()
            if best is None or c["kernel_ms"] < best["kernel_ms"]:
                best = c
        n = res * res * spp
        print(f"{name:10s} {str(opts):36s} slots {kl.getOption('warp_slots')} edge {kl.getOption('skip'):>3s} shape {kl.launchShape()} "
              f"{best['kernel_ms']:8.3f} ms {n / best['kernel_ms'] / 1e3:8.1f} Msamples/s  lookups/path {best['density_lookups'] / n:6.1f} "
              f"skipped {best['skipped_fetches'] / max(best['density_lookups'], 1):.3f}", flush=True)
        kl.close()
