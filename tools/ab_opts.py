"""A/B of arbitrary option sets on the usual scenes: kernel ms, Msamples/s.
usage: python tools/ab_opts.py res spp "k=v,k=v" "k=v" ...   (an empty string = defaults)
       CVR_AB_SCENES=hetvol,manix restricts the scenes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudavolumerenderer_b200 import createLauncher, scenes
from cudavolumerenderer_b200.launcher import ProceduralScene

res, spp = int(sys.argv[1]), int(sys.argv[2])
sets = [dict(kv.split("=", 1) for kv in a.split(",") if kv) for a in sys.argv[3:]] or [{}]
CASES = [("bucky", lambda: scenes.bucky(), "naiveSK"), ("hetvol", lambda: scenes.hetvol(), "regenerationSK"),
         ("manix", lambda: scenes.manix(), "regenerationSK"), ("fbm512", lambda: ProceduralScene("fbm", 512), "regenerationSK"),
         ("fbm1024", lambda: ProceduralScene("fbm", 1024), "regenerationSK"),
         ("sparse1024", lambda: ProceduralScene("sparsefbm", 1024), "regenerationSK")]
only = os.environ.get("CVR_AB_SCENES")
for name, make, kernel in CASES:
    if only and name not in only.split(","):
        continue
    sc = make()
    for opts in sets:
        o = dict(opts)
        k = o.pop("kernel", kernel)
        try:
            kl = createLauncher(k, 0, **o)
            kl.setScene(sc)
            best = None
            for rep in range(3):
                kl.resetCounters()
                kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
                c = kl.counters()
                if best is None or c["kernel_ms"] < best["kernel_ms"]:
                    best = c
            n = res * res * spp
            print(f"{name:10s} {str(opts):60s} slots {kl.getOption('warp_slots')} skip {kl.getOption('skip'):>3s} shape {kl.launchShape()} "
                  f"{best['kernel_ms']:8.3f} ms {n / best['kernel_ms'] / 1e3:8.1f} Msamples/s  lookups/path {best['density_lookups'] / n:6.1f} "
                  f"skipped {best['skipped_fetches'] / max(best['density_lookups'] + best['speculative_lookups'], 1):.3f}", flush=True)
            kl.close()
        except Exception as e:
            print(f"{name:10s} {str(opts):60s} FAILED {e}", flush=True)
