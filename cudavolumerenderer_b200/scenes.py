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


def load(path: str, scene_type: str = "Auto", name: str | None = None) -> Scene:
    """SceneAssembler::getScene over the C++ SceneBuilders (cvr_scene_file_load, include/cvr_abi.h): the ONE
    implementation of the reference's loaders (RawSceneBuilder.h:35-140, XmlSceneBuilder.h:39-266,
    VDBSceneBuilder.h:40-80) and of the stand-in parameter tables above; this module only wraps the result.
    scene_type as ConfigParser.cpp:84-103: "Auto" (by extension) | "Raw" | "MitsubaXml" | "Vdb"."""
    d = abi.load_scene_file(path, scene_type)
    return Scene(d["density"], d["albedo"], d["box_min"], d["box_max"], scale=d["scale"], max_density=d["max_density"],
                 fov_x=d["fov_x"], albedo_const=d["albedo_const"], hg_g=d["hg_g"], ggx_alpha=d["ggx_alpha"],
                 ggx_eta=d["ggx_eta"], name=name or path.rsplit("/", 1)[-1])


def _synth(name: str, dims=None, seed: int = 0) -> Scene:
    spec = f"synth:{name}"
    if dims is not None:
        spec += ":" + "x".join(str(int(v)) for v in dims)
    if seed:
        spec += f":seed={int(seed)}"
    return load(spec, name=name)


def bucky(seed: int = 0) -> Scene:
    return _synth("bucky", None, seed)


def hetvol(seed: int = 0, dims=(128, 128, 50)) -> Scene:
    return _synth("hetvol", dims, seed)


def manix(seed: int = 0, dims=(256, 230, 256)) -> Scene:
    return _synth("manix", dims, seed)


def fbm(n: int = 256, seed: int = 0, albedo: float = 0.99) -> Scene:
    sc = _synth("fbm", (n, n, n), seed)
    sc.albedo_const = (float(albedo),) * 3
    sc.name = f"fbm{n}"
    return sc


def from_vdb(path: str) -> Scene:
    """VDBSceneBuilder (VDBSceneBuilder.h:40-80): density FloatGrid + albedo Vec3SGrid densified
    over the density grid's active bounding box, max_density = max voxel, the file's world
    box read and ignored (box fixed to +-0.5, Q4), scale 100, default camera (fov 0.7)."""
    return load(path, "Vdb")


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
