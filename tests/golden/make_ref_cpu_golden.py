"""Generate tests/golden/ref_cpu_paths.npz from the REFERENCE's own path kernels compiled
for the host (oracle/_ref/libcvr_ref_cpu.so = NaiveVolPTsk_kernel::d_render and
RegenerationVolPTsk_kernel::d_render_single_thread_regeneration built by g++ from the
headers under /root/reference, oracle/ref_cpu_harness.cpp).
Run in the build container:  python tests/golden/make_ref_cpu_golden.py

Per-path radiances (naiveSK, Rng(path id)) and single-stream regenerationSK images on the
repo's deterministic procedural scenes; they pin oracle/cvr_oracle.c's WHOLE path loop
(camera, box test, Woodcock loop, trilinear lookups, GGX boundary, HG scatter, roulette)
bit for bit on boxes where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from cudavolumerenderer_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, scene, tile (w, h), full (w, h), tile offset, naive spp, regen spp, regen seed
CASES = [
    ("bucky", "bucky", (32, 32), (32, 32), (0, 0), 4, 2, 0),
    ("hetvol", "hetvol", (32, 32), (32, 32), (0, 0), 2, 1, 12345),
    # a tile of a larger image: offset + pixel_index_range differ from the tile size (A4, A15)
    ("bucky_tile", "bucky", (24, 16), (96, 64), (40, 24), 4, 2, 7),
]


def case_inputs(c):
    name, scene, tile, full, off, spp, rspp, rseed = c
    sc = scenes.make(scene)
    osc = B.make_scene(sc.density, sc.albedo, sc.box_min, sc.box_max, sc.scale, sc.max_density)
    cam = B.make_camera(tile[0], tile[1], full[0], full[1], off[0], off[1], fov_x=sc.fov_x)
    return sc, osc, cam


def main():
    out = {}
    for c in CASES:
        name, scene, tile, full, off, spp, rspp, rseed = c
        sc, osc, cam = case_inputs(c)
        rc = B.RefCpu(osc, cam)
        n = tile[0] * tile[1] * spp
        out[name + "_paths"] = rc.trace_paths_naive(0, n, n_threads=1)
        out[name + "_regen"] = rc.render_regen(rspp, seed=rseed, n_threads=1)
        print(name, n, "paths, mean", float(out[name + "_paths"][:, :3].mean()),
              "escaped", int((out[name + "_paths"][:, 3] == 1).sum()))
    np.savez_compressed(os.path.join(HERE, "ref_cpu_paths.npz"), **out)


if __name__ == "__main__":
    main()
