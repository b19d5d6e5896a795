"""ctypes binding of libcvr_b200.so (include/cvr_abi.h).

This is plumbing for tests/bench and for Python users; the product is the C-ABI
library.  There is NO fallback: if the library is missing or a call fails, an
exception is raised (``CvrError``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVR_B200_LIB", os.path.join(HERE, "libcvr_b200.so"))  # override: kernel-tuning experiments only
# experiments only (A/B of two builds in one GPU call): CVR_LIB=<path to another build of the same ABI>
LIB_PATH = os.environ.get("CVR_LIB", LIB_PATH)

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


class CvrError(RuntimeError):
    pass


class SceneDesc(C.Structure):
    _fields_ = [
        ("density", C.c_void_p), ("density_dim", C.c_int32 * 3),
        ("albedo", C.c_void_p), ("albedo_dim", C.c_int32 * 3),
        ("albedo_const", C.c_float * 3),
        ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
        ("scale", C.c_float), ("max_density", C.c_float), ("hg_g", C.c_float),
        ("ggx_alpha", C.c_float * 2), ("ggx_eta", C.c_float),
        ("density_on_device", C.c_int32),
    ]


class SceneFileInfo(C.Structure):  # cvr_scene_file_info_t
    _fields_ = [
        ("scene", SceneDesc), ("resolution", C.c_uint32 * 2), ("fov_x", C.c_float), ("inv_view", C.c_float * 12),
        ("raster_to_view", C.c_float * 2), ("type", C.c_char * 16),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("bounces", C.c_uint64), ("density_lookups", C.c_uint64),
        ("albedo_lookups", C.c_uint64), ("escaped", C.c_uint64),
        ("speculative_lookups", C.c_uint64), ("launches", C.c_uint64), ("kernel_ms", C.c_double),
        ("skipped_fetches", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SparseDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32 * 3), ("bbox_min", C.c_int32 * 3), ("n_leaves", C.c_uint64),
        ("leaf_origins", C.c_void_p), ("leaf_values", C.c_void_p),
        ("albedo_const", C.c_float * 3), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
        ("scale", C.c_float), ("max_density", C.c_float), ("hg_g", C.c_float),
        ("ggx_alpha", C.c_float * 2), ("ggx_eta", C.c_float),
    ]


class VdbGridInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 64), ("type", C.c_char * 64), ("channels", C.c_int32), ("compression", C.c_uint32),
        ("file_version", C.c_uint32), ("bbox_min", C.c_int32 * 3), ("bbox_max", C.c_int32 * 3), ("dim", C.c_int32 * 3),
        ("background", C.c_float * 3), ("active_voxels", C.c_uint64), ("leaf_count", C.c_uint64),
        ("active_tiles", C.c_uint64),
    ]


class RenderDesc(C.Structure):
    _fields_ = [
        ("res_x", C.c_uint32), ("res_y", C.c_uint32),
        ("n_tiles_x", C.c_uint32), ("n_tiles_y", C.c_uint32),
        ("iterations", C.c_uint32), ("fov_x", C.c_float),
        ("inv_view", f32p), ("raster_to_view", f32p),
        ("tile_first", C.c_uint32), ("tile_stride", C.c_uint32),
        ("sample_first", C.c_uint32), ("sample_count", C.c_uint32),
        ("fuse_tiles", C.c_int32),
    ]


class Shard(C.Structure):
    """cvr_shard: one rank's share of a render (whole tiles + sample range of the tail tiles)."""
    _fields_ = [("tile_first", C.c_uint32), ("tile_stride", C.c_uint32), ("tile_limit", C.c_uint32),
                ("tail_first", C.c_uint32), ("tail_limit", C.c_uint32),
                ("sample_first", C.c_uint32), ("sample_count", C.c_uint32)]


SHARD_MODES = {"tiles": 0, "spp": 1, "balanced": 2}

# every symbol include/cvr_abi.h declares: (name, restype, argtypes)
H = C.c_void_p
SYMBOLS = [
    ("cvr_create", C.c_int, [C.c_char_p, C.c_int, C.POINTER(H)]),
    ("cvr_destroy", C.c_int, [H]),
    ("cvr_last_error", C.c_char_p, [H]),
    ("cvr_abi_version", C.c_int, []),
    ("cvr_set_option", C.c_int, [H, C.c_char_p, C.c_char_p]),
    ("cvr_get_option", C.c_int, [H, C.c_char_p, C.c_char_p, C.c_size_t]),
    ("cvr_set_stream", C.c_int, [H, C.c_void_p]),
    ("cvr_set_scene", C.c_int, [H, C.POINTER(SceneDesc)]),
    ("cvr_set_resolution", C.c_int, [H, C.c_uint32, C.c_uint32]),
    ("cvr_set_pixel_index_range", C.c_int, [H, C.c_float, C.c_float]),
    ("cvr_set_raster_to_view", C.c_int, [H, C.c_float, C.c_float]),
    ("cvr_set_inv_view_matrix", C.c_int, [H, f32p]),
    ("cvr_set_offset", C.c_int, [H, C.c_uint32, C.c_uint32]),
    ("cvr_set_output", C.c_int, [H, C.c_void_p]),
    ("cvr_set_iterations", C.c_int, [H, C.c_uint32]),
    ("cvr_get_iterations", C.c_int, [H, u32p]),
    ("cvr_set_seed", C.c_int, [H, C.c_uint32]),
    ("cvr_get_seed", C.c_int, [H, u32p]),
    ("cvr_set_sample_range", C.c_int, [H, C.c_uint32, C.c_uint32]),
    ("cvr_init", C.c_int, [H]),
    ("cvr_allocate", C.c_int, [H]),
    ("cvr_launch_render", C.c_int, [H]),
    ("cvr_reset", C.c_int, [H]),
    ("cvr_sync", C.c_int, [H]),
    ("cvr_release", C.c_int, [H]),
    ("cvr_get_counters", C.c_int, [H, C.POINTER(Counters)]),
    ("cvr_reset_counters", C.c_int, [H]),
    ("cvr_get_launch_shape", C.c_int, [H, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("cvr_resolve_tile", C.c_int, [H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]),
    ("cvr_resolve_tile_display", C.c_int, [H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_float, C.c_int]),
    ("cvr_render_image", C.c_int, [H, C.POINTER(RenderDesc), C.c_void_p, C.c_void_p]),
    ("cvr_shard_plan", C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(Shard)]),
    ("cvr_render_image_sharded", C.c_int, [H, C.POINTER(RenderDesc), C.POINTER(Shard), C.c_void_p, C.c_void_p]),
    ("cvr_group_create", C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(H)]),
    ("cvr_group_destroy", C.c_int, [H]),
    ("cvr_group_last_error", C.c_char_p, [H]),
    ("cvr_group_size", C.c_int, [H, C.POINTER(C.c_int)]),
    ("cvr_group_member", C.c_int, [H, C.c_int, C.POINTER(H)]),
    ("cvr_group_set_option", C.c_int, [H, C.c_char_p, C.c_char_p]),
    ("cvr_group_set_seed", C.c_int, [H, C.c_uint32]),
    ("cvr_group_set_scene", C.c_int, [H, C.POINTER(SceneDesc)]),
    ("cvr_group_set_scene_sparse", C.c_int, [H, C.c_void_p]),
    ("cvr_group_set_scene_procedural", C.c_int, [H, C.c_char_p, C.c_int32, C.c_uint32, C.c_void_p, f32p]),
    ("cvr_group_render_image", C.c_int, [H, C.POINTER(RenderDesc), C.c_int, C.c_void_p, C.c_void_p]),
    ("cvr_group_reduce", C.c_int, [H, C.POINTER(C.c_void_p), C.c_uint64]),
    ("cvr_group_get_counters", C.c_int, [H, C.POINTER(Counters)]),
    ("cvr_group_reset_counters", C.c_int, [H]),
    ("cvr_tile_table", C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p]),
    ("cvr_default_camera", C.c_int, [C.c_uint32, C.c_uint32, C.c_float, f32p, f32p]),
    ("cvr_trace_paths", C.c_int, [H, C.c_uint64, C.c_uint64, C.c_void_p]),
    ("cvr_trace_paths_logged", C.c_int, [H, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]),
    ("cvr_rng_kat", C.c_int, [H, C.POINTER(C.c_int32), C.c_int, C.c_int, u32p, f32p]),
    ("cvr_debug_lookup", C.c_int, [H, f32p, C.c_int, f32p, f32p]),
    ("cvr_debug_trig_check", C.c_int, [H, C.c_float, C.POINTER(C.c_uint64), u32p]),
    ("cvr_gather_roofline", C.c_int, [H, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("cvr_synth_volume", C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, f32p, f32p, f32p]),
    ("cvr_set_scene_sparse", C.c_int, [C.c_void_p, C.c_void_p]),
    ("cvr_set_scene_procedural", C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_uint32, C.c_void_p, f32p]),
    ("cvr_get_volume_info", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    ("cvr_scene_file_load", C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    ("cvr_scene_file_info", C.c_int, [C.c_void_p, C.POINTER(SceneFileInfo)]),
    ("cvr_scene_file_close", C.c_int, [C.c_void_p]),
    ("cvr_scene_file_last_error", C.c_char_p, []),
    ("cvr_vdb_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    ("cvr_vdb_close", C.c_int, [C.c_void_p]),
    ("cvr_vdb_last_error", C.c_char_p, []),
    ("cvr_vdb_grid_count", C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("cvr_vdb_grid_info", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    ("cvr_vdb_grid_meta", C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    ("cvr_vdb_densify", C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, f32p, f32p, C.c_uint64]),
    ("cvr_vdb_leaves", C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_int32),
                                 C.POINTER(C.c_uint64), f32p]),
]

_lib = None


def load() -> C.CDLL:
    """Load libcvr_b200.so and bind every symbol; raises CvrError when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CvrError(
                f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cudavolumerenderer_b200/csrc`); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(h, rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().cvr_last_error(h)
        raise CvrError(f"{what}: {msg.decode() if msg else 'error'}")


def tile_table(res_x: int, res_y: int, ntx: int, nty: int):
    dim = np.zeros(2, np.uint32)
    org = np.zeros((ntx * nty, 2), np.uint32)
    rc = load().cvr_tile_table(res_x, res_y, ntx, nty, dim.ctypes.data_as(u32p), org.ctypes.data_as(u32p))
    if rc:
        raise CvrError("cvr_tile_table failed")
    return dim, org


def shard_plan(n_tiles: int, iterations: int, rank: int, world: int, mode: str = "balanced") -> Shard:
    """cvr_shard_plan: rank `rank`'s share of n_tiles tiles x iterations samples (mode tiles | spp | balanced)."""
    sh = Shard()
    if mode not in SHARD_MODES or load().cvr_shard_plan(n_tiles, iterations, rank, world, SHARD_MODES[mode], C.byref(sh)):
        raise ValueError(f"bad shard plan request: tiles {n_tiles} iterations {iterations} rank {rank}/{world} mode {mode}")
    return sh


def default_camera(res_x: int, res_y: int, fov_x: float = 0.7):
    iv = np.zeros(12, np.float32)
    rtv = np.zeros(2, np.float32)
    load().cvr_default_camera(res_x, res_y, fov_x, iv.ctypes.data_as(f32p), rtv.ctypes.data_as(f32p))
    return iv, rtv


def load_scene_file(path: str, scene_type: str = "Auto"):
    """cvr_scene_file_load + cvr_scene_file_info: the C++ SceneBuilders of the host layer (host/SceneBuilders.h).
    Returns a dict of numpy COPIES of the volumes and the medium / camera parameters."""
    lib = load()
    h = C.c_void_p()
    if lib.cvr_scene_file_load(path.encode(), scene_type.encode(), C.byref(h)):
        raise CvrError(lib.cvr_scene_file_last_error().decode())
    try:
        info = SceneFileInfo()
        if lib.cvr_scene_file_info(h, C.byref(info)):
            raise CvrError(lib.cvr_scene_file_last_error().decode())
        d = info.scene
        nx, ny, nz = d.density_dim
        den = np.ctypeslib.as_array(C.cast(d.density, f32p), shape=(nz, ny, nx)).copy()
        alb = None
        if d.albedo:
            ax, ay, az = d.albedo_dim
            alb = np.ctypeslib.as_array(C.cast(d.albedo, f32p), shape=(az, ay, ax, 4)).copy()
        return {"density": den, "albedo": alb, "albedo_const": tuple(d.albedo_const), "box_min": tuple(d.box_min),
                "box_max": tuple(d.box_max), "scale": float(d.scale), "max_density": float(d.max_density),
                "hg_g": float(d.hg_g), "ggx_alpha": tuple(d.ggx_alpha), "ggx_eta": float(d.ggx_eta),
                "resolution": tuple(info.resolution), "fov_x": float(info.fov_x),
                "inv_view": np.array(info.inv_view, np.float32), "type": info.type.decode()}
    finally:
        lib.cvr_scene_file_close(h)


def synth_volume(kind: str, nx: int, ny: int, nz: int, seed: int = 0, with_albedo: bool = True):
    """Returns (density (nz,ny,nx) f32, albedo (nz,ny,nx,4) f32 or None, max_density)."""
    den = np.zeros((nz, ny, nx), np.float32)
    alb = np.zeros((nz, ny, nx, 4), np.float32) if with_albedo else None
    mx = C.c_float()
    rc = load().cvr_synth_volume(kind.encode(), nx, ny, nz, seed, den.ctypes.data_as(f32p),
                                 alb.ctypes.data_as(f32p) if with_albedo else None, C.byref(mx))
    if rc:
        raise CvrError(f"cvr_synth_volume({kind}) failed")
    return den, alb, float(mx.value)
