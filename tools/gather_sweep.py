"""Random-sector gather microbenchmark over footprints, L1 allowed / bypassed, per L2 fetch granularity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudavolumerenderer_b200 import RegenerationVolPTsk
kl = RegenerationVolPTsk(0)
for fetch in ("default", "32", "64", "128"):
    kl.setOption("l2_fetch", fetch)
    print("l2_fetch", fetch, "->", kl.getOption("l2_fetch"))
    for fp in (16 << 20, 104 << 20, 512 << 20, 4 << 30, 32 << 30):
        a = kl.gatherRoofline(fp, 512, 8)
        b = kl.gatherRoofline(fp, 512, 8, bypass_l1=True)
        print(f"  footprint {fp / 2**20:9.0f} MiB  L1 allowed {a:8.1f} GB/s   L1 bypassed {b:8.1f} GB/s", flush=True)
kl.setOption("l2_fetch", "default")
kl.close()
