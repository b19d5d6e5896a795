"""Small driver for ncu captures: renders one workload a few times through the C ABI.
usage: python tools/profile_run.py [scene] [res] [spp] [reps] [key=value options...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudavolumerenderer_b200 import scenes, createLauncher

scene = sys.argv[1] if len(sys.argv) > 1 else "hetvol"
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
opts = dict(a.split("=", 1) for a in sys.argv[5:])
kernel = opts.pop("kernel", "regenerationSK")
if scene.startswith("dev"):  # devfbm:1024 | devsparsefbm:2048 -> generated on the device
    from cudavolumerenderer_b200.launcher import ProceduralScene
    sc = ProceduralScene(scene.split(':')[0][3:], int(scene.split(':')[1]))
else:
    sc = scenes.make(scene.split(':')[0], **({'n': int(scene.split(':')[1])} if ':' in scene else {}))
kl = createLauncher(kernel, 0, **opts)
kl.setScene(sc)
for i in range(reps):
    kl.resetCounters()
    t0 = time.perf_counter()
    img = kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
    dt = time.perf_counter() - t0
    c = kl.counters()
    n = res * res * spp
    alg = 32 * c["density_lookups"] + (128 * c["albedo_lookups"] if getattr(sc, "albedo", None) is not None else 0) + 16 * c["paths"]
    print(f"{scene} {res}x{res}x{spp} {kernel} {opts}: kernel {c['kernel_ms']:.3f} ms  {n / c['kernel_ms'] / 1e3:.1f} Msamples/s  "
          f"{c['density_lookups'] / c['kernel_ms'] / 1e6:.2f} Glookups/s  alg {alg / c['kernel_ms'] / 1e6:.0f} GB/s  "
          f"lookups/path {c['density_lookups'] / n:.1f} bounces/path {c['bounces'] / n:.2f} albedo/path {c['albedo_lookups'] / n:.2f} "
          f"nan_px {int(np.isnan(img).any(axis=2).sum())} mean {np.nanmean(img[..., :3]):.4f} wall {dt * 1e3:.1f} ms shape {kl.launchShape()}")
kl.close()
