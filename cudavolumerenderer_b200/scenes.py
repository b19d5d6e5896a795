"""Named scenes of BASELINE.json's configs, as procedural stand-ins.

Every volume payload in the reference checkout is a Git-LFS pointer stub, so the
scenes are synthesised by libcvr_b200.so (csrc/cvr_synth.cpp) with the grid shapes and
medium parameters the reference's loaders would produce:

  bucky   RawSceneBuilder.h:35-83     32^3, box +-0.5, scale 40, max_density 1, fov 0.7
  hetvol  XmlSceneBuilder.h:39-152    density 128x128x50, albedo grid box
          [-0.64,-0.64,-0.25]..[0.64,0.64,0.25] becomes the medium box (Q3), scale 800,
          fov 0.33 (data/mitsubaxml/smoke/hetvol.xml)
  manix   VDBSceneBuilder.h:40-80     256x230x256, box +-0.5, scale 100, fov 0.7
  fbm     synthetic n^3, constant albedo 0.99, scale 100
"""
from __future__ import annotations

from . import abi
from .launcher import Scene


def bucky(seed: int = 0) -> Scene:
    den, alb, mx = abi.synth_volume("bucky", 32, 32, 32, seed)
    return Scene(den, alb, (-0.5,) * 3, (0.5,) * 3, scale=40.0, max_density=mx, fov_x=0.7, name="bucky")


def hetvol(seed: int = 0, dims=(128, 128, 50)) -> Scene:
    den, alb, mx = abi.synth_volume("hetvol", dims[0], dims[1], dims[2], seed)
    return Scene(den, alb, (-0.64, -0.64, -0.25), (0.64, 0.64, 0.25), scale=800.0, max_density=mx,
                 fov_x=0.33, name="hetvol")


def manix(seed: int = 0, dims=(256, 230, 256)) -> Scene:
    den, alb, mx = abi.synth_volume("manix", dims[0], dims[1], dims[2], seed)
    return Scene(den, alb, (-0.5,) * 3, (0.5,) * 3, scale=100.0, max_density=mx, fov_x=0.7, name="manix")


def fbm(n: int = 256, seed: int = 0, albedo: float = 0.99) -> Scene:
    den, _, mx = abi.synth_volume("fbm", n, n, n, seed, with_albedo=False)
    return Scene(den, None, (-0.5,) * 3, (0.5,) * 3, scale=100.0, max_density=mx, fov_x=0.7,
                 albedo_const=(albedo,) * 3, name=f"fbm{n}")


def from_vdb(path: str) -> Scene:
    """VDBSceneBuilder (VDBSceneBuilder.h:40-80): density FloatGrid + albedo Vec3SGrid densified
    over the density grid's active bounding box, max_density = max voxel, the file's world
    box read and ignored (box fixed to +-0.5, Q4), scale 100, default camera (fov 0.7)."""
    from .vdb import VdbFile

    with VdbFile(path) as f:
        den = f.densify("density", inactive=(0.0, 0.0, 0.0))
        alb_grid = f.grid("albedo")  # raises "VDB file does not contain an albedo grid"
        if alb_grid["dim"] != f.grid("density")["dim"]:
            # the reference indexes the albedo array with the ALBEDO box but sizes the volume with
            # the DENSITY resolution (VDBSceneBuilder.h:47-67): only defined when they agree
            raise abi.CvrError("density and albedo grids have different active bounding boxes")
        alb = f.densify("albedo", out_channels=4, inactive=(0.0, 0.0, 0.0))
    return Scene(den, alb, (-0.5,) * 3, (0.5,) * 3, scale=100.0, max_density=float(den.max()), fov_x=0.7,
                 name=path.rsplit("/", 1)[-1])


def sparse_from_vdb(path: str, albedo_const=(1.0, 1.0, 1.0)):
    """The density grid of a .vdb file as a SparseScene: its leaves go to device bricks without
    densifying (lookups identical to from_vdb's dense grid; constant albedo)."""
    from .launcher import SparseScene
    from .vdb import VdbFile
    import numpy as np

    with VdbFile(path) as f:
        g = f.grid("density")
        org, msk, val = f.leaves("density")
    bits = np.unpackbits(msk.view(np.uint8).reshape(len(msk), 64), axis=1, bitorder="little").astype(bool)
    val = np.where(bits, val, np.float32(0.0))  # inactive voxels read as 0 (VDBAdapter.cpp:66)
    return SparseScene(org, val, g["dim"], g["bbox_min"], scale=100.0, albedo_const=albedo_const,
                       name=path.rsplit("/", 1)[-1])


def fbm_device(n: int = 1024, seed: int = 0, albedo: float = 0.99):
    """C4: fBm n^3 generated on the device into the dense cell8 layout (no host staging)."""
    from .launcher import ProceduralScene

    return ProceduralScene("fbm", n, seed, albedo_const=(albedo,) * 3)


def sparse_fbm(n: int = 2048, seed: int = 0, albedo: float = 0.99):
    """C5: VDB-style sparse n^3 volume (~3 % of the 8^3 bricks active) generated on the device
    into the brick layout."""
    from .launcher import ProceduralScene

    return ProceduralScene("sparsefbm", n, seed, albedo_const=(albedo,) * 3)


SCENES = {"bucky": bucky, "hetvol": hetvol, "manix": manix, "fbm": fbm}


def make(name: str, **kw) -> Scene:
    if name not in SCENES:
        raise ValueError(f"unknown scene '{name}' (choices: {sorted(SCENES)})")
    return SCENES[name](**kw)
