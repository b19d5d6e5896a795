"""Per-rank kernel time of a sharded C3 render, measured on ONE GPU: rank r's share of the plan is rendered alone and
timed (CUDA events the library records around its launches), for every rank of a world of N.  Shows how much of the
strong-scaling loss is work imbalance between the shares and how much is fixed cost per launch.
usage: python tools/shard_balance.py [world=8] [modes=tiles,spp,balanced]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cudavolumerenderer_b200 import RegenerationVolPTsk, abi, scenes

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
modes = (sys.argv[2] if len(sys.argv) > 2 else "tiles,spp,balanced").split(",")
RES, SPP, TILES = 1024, 256, (10, 10)
sc = scenes.manix()
kl = RegenerationVolPTsk(0)
kl.setScene(sc)
d_img = torch.zeros((RES, RES, 4), dtype=torch.float32, device="cuda:0")
full = None
for rep in range(3):
    kl.resetCounters()
    kl.setSeed(0)
    kl.renderImage((RES, RES), TILES, SPP, fov_x=sc.fov_x, fuse_tiles=True, d_image=d_img.data_ptr())
    torch.cuda.synchronize()
    c = kl.counters()
    full = c["kernel_ms"] if full is None else min(full, c["kernel_ms"])
print(f"one GPU, whole image: kernel {full:.3f} ms ({c['launches']} launch); ideal share of {world}: {full / world:.3f} ms")
for mode in modes:
    times, launches = [], []
    for r in range(world):
        sh = abi.shard_plan(TILES[0] * TILES[1], SPP, r, world, mode)
        best = None
        for rep in range(2):
            kl.resetCounters()
            kl.setSeed(0)
            kl.renderImageSharded((RES, RES), TILES, SPP, sh, fov_x=sc.fov_x, d_image=d_img.data_ptr())
            torch.cuda.synchronize()
            c = kl.counters()
            best = c["kernel_ms"] if best is None else min(best, c["kernel_ms"])
        times.append(best)
        launches.append(int(c["launches"]))
    print(f"{mode:9s} world {world}: per-rank kernel ms " + " ".join(f"{t:.3f}" for t in times) +
          f" | max {max(times):.3f} mean {sum(times) / world:.3f} sum {sum(times):.3f} launches/rank {launches[0]}..{launches[-1]}"
          f" | efficiency bound {full / world / max(times):.3f}")
kl.close()
