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
# scene = a name of cudavolumerenderer_b200.scenes or (name, kwargs)
CASES = [
    ("bucky", "bucky", (32, 32), (32, 32), (0, 0), 4, 2, 0),
    ("hetvol", "hetvol", (32, 32), (32, 32), (0, 0), 2, 1, 12345),
    # a tile of a larger image: offset + pixel_index_range differ from the tile size (A4, A15)
    ("bucky_tile", "bucky", (24, 16), (96, 64), (40, 24), 4, 2, 7),
    # C3: MANIX-like head phantom, albedo = (rho, 0, 0): a reduced grid on a full tile ...
    ("manix", ("manix", {"dims": (96, 80, 72)}), (32, 32), (32, 32), (0, 0), 4, 2, 5),
    # ... and the FULL 256x230x256 grid seen through a 64^2 tile of the 1024^2 north-star image
    ("manix_c3_tile", "manix", (64, 64), (1024, 1024), (480, 448), 2, 0, 0),
    # C4: fBm with constant albedo 0.99 (long multi-scatter paths)
    ("fbm", ("fbm", {"n": 64}), (32, 32), (32, 32), (0, 0), 2, 1, 9),
    # C5: VDB-style sparse volume (~8 % of the bricks active), densified for the reference kernels
    ("sparsefbm", ("sparsefbm", {"n": 96, "seed": 4}), (32, 32), (32, 32), (0, 0), 2, 1, 3),
]


def make_scene(spec):
    """The host scene of a case: a cudavolumerenderer_b200 Scene whose albedo is an ARRAY (the
    reference kernels and the oracle take volumes only: a constant albedo becomes a 2x2x2 grid)."""
    from cudavolumerenderer_b200 import Scene, abi

    name, kw = (spec, {}) if isinstance(spec, str) else spec
    if name == "sparsefbm":
        n, seed = kw["n"], kw.get("seed", 0)
        den, _, mx = abi.synth_volume("sparsefbm", n, n, n, seed, with_albedo=False)
        sc = Scene(den, None, (-0.5,) * 3, (0.5,) * 3, scale=100.0, max_density=mx, albedo_const=(0.99,) * 3,
                   name=f"sparsefbm{n}")
    else:
        sc = scenes.make(name, **kw)
    if sc.albedo is None:
        alb = np.empty((2, 2, 2, 4), np.float32)
        alb[..., :3] = np.asarray(sc.albedo_const, np.float32)
        alb[..., 3] = 1.0
        sc.albedo_array = alb
    else:
        sc.albedo_array = sc.albedo
    return sc


def case_inputs(c):
    name, scene, tile, full, off, spp, rspp, rseed = c
    sc = make_scene(scene)
    osc = B.make_scene(sc.density, sc.albedo_array, sc.box_min, sc.box_max, sc.scale, sc.max_density)
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
        if rspp:
            out[name + "_regen"] = rc.render_regen(rspp, seed=rseed, n_threads=1)
        print(name, n, "paths, mean", float(out[name + "_paths"][:, :3].mean()),
              "escaped", int((out[name + "_paths"][:, 3] == 1).sum()))
    np.savez_compressed(os.path.join(HERE, "ref_cpu_paths.npz"), **out)


if __name__ == "__main__":
    main()
