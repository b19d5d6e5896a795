"""ctypes bindings for the CPU oracle (oracle/libcvr_oracle.so) and, when built,
the reference-derived checkers under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(cudavolumerenderer_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libcvr_oracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libcvr_ref_host.so")
REF_GPU_SO = os.path.join(HERE, "_ref", "libcvr_ref_gpu.so")
REF_CPU_SO = os.path.join(HERE, "_ref", "libcvr_ref_cpu.so")

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


class Rng(C.Structure):
    _fields_ = [("v", C.c_uint32 * 5), ("d", C.c_uint32)]


class Scene(C.Structure):
    _fields_ = [
        ("density", f32p), ("dnx", C.c_int32), ("dny", C.c_int32), ("dnz", C.c_int32),
        ("albedo", f32p), ("anx", C.c_int32), ("any", C.c_int32), ("anz", C.c_int32),
        ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
        ("scale", C.c_float), ("max_density", C.c_float), ("hg_g", C.c_float),
        ("ggx_alpha", C.c_float * 2), ("ggx_eta", C.c_float),
    ]


class Camera(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 12), ("raster_to_view", C.c_float * 2),
        ("resolution", C.c_float * 2), ("pixel_index_range", C.c_float * 2),
        ("offset", C.c_uint32 * 2),
    ]


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("bounces", C.c_uint64),
                ("density_lookups", C.c_uint64), ("albedo_lookups", C.c_uint64),
                ("escaped", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Trace(C.Structure):
    """cvro_trace (cvr_oracle.h): event / decision log of one path."""
    _fields_ = [("d0", C.c_uint32), ("n_events", C.c_uint32), ("cap_events", C.c_uint32),
                ("ev_code", u32p), ("ev_d", u32p),
                ("n_dec", C.c_uint32), ("cap_dec", C.c_uint32),
                ("dec_kind", u32p), ("dec_d", u32p), ("dec_a", f32p), ("dec_b", f32p)]


EV_SCATTER, EV_BOUNDARY, EV_ESCAPE = 1, 2, 3
EVF_OK, EVF_WO_NEG, EVF_WI_NEG, EVF_KILLED, EVF_ZERO = 16, 32, 64, 128, 256
DEC_EXIT, DEC_ACCEPT, DEC_ROULETTE, DEC_FRESNEL = 1, 2, 3, 4
XORWOW_D_STEP = 362437  # the draw counter `d` advances by this per draw


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.exists(ORACLE_SO) or (
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "cvr_oracle.c"))):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/implementation/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.cvro_rng_init.argtypes = [C.POINTER(Rng), C.c_int32]
        L.cvro_rng_u32.argtypes = [C.POINTER(Rng)]
        L.cvro_rng_u32.restype = C.c_uint32
        L.cvro_rng_float.argtypes = [C.POINTER(Rng)]
        L.cvro_rng_float.restype = C.c_float
        L.cvro_trace_path.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.POINTER(Rng),
                                      C.c_uint32, C.c_int, C.c_uint32, f32p, C.POINTER(Counters)]
        L.cvro_trace_path.restype = C.c_int
        L.cvro_render_naive.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.c_uint32, f32p,
                                        C.c_int, C.POINTER(Counters)]
        L.cvro_render_naive.restype = None
        L.cvro_trace_paths_naive.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.c_uint64,
                                             C.c_uint64, f32p, C.c_int, C.POINTER(Counters)]
        L.cvro_trace_paths_naive.restype = None
        L.cvro_trace_paths_seeded.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.c_uint64, C.c_uint64,
                                              C.c_uint32, C.c_int, f32p, C.c_int, C.POINTER(Counters)]
        L.cvro_trace_paths_seeded.restype = None
        L.cvro_trace_path_logged.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.c_int32, C.c_uint32, C.c_int,
                                             C.c_uint32, f32p, C.POINTER(Trace)]
        L.cvro_trace_path_logged.restype = C.c_int
        L.cvro_render_regen.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.c_uint32,
                                        C.c_uint32, C.c_int, C.c_uint32, f32p, C.c_int,
                                        C.POINTER(Counters)]
        L.cvro_render_regen.restype = None
        L.cvro_aabb_intersect.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.POINTER(C.c_int)]
        L.cvro_aabb_intersect.restype = C.c_int
        L.cvro_frame_from_z.argtypes = [f32p] * 4
        L.cvro_hg_sample.argtypes = [f32p, C.c_float, C.c_float, C.c_float, f32p]
        L.cvro_ggx_sample.argtypes = [f32p, C.c_float, f32p, f32p, f32p, f32p, C.POINTER(C.c_int)]
        L.cvro_ggx_sample.restype = C.c_int
        L.cvro_fresnel_dielectric.argtypes = [C.c_float, C.c_float, f32p]
        L.cvro_fresnel_dielectric.restype = C.c_float
        L.cvro_ggx_g1.argtypes = [f32p, f32p, f32p]
        L.cvro_ggx_g1.restype = C.c_float
        L.cvro_density_lookup.argtypes = [C.POINTER(Scene), f32p]
        L.cvro_density_lookup.restype = C.c_float
        L.cvro_woodcock.argtypes = [C.POINTER(Scene), f32p, f32p, C.c_float, C.c_int32, C.POINTER(C.c_int)]
        L.cvro_woodcock.restype = C.c_float
        L.cvro_albedo_lookup.argtypes = [C.POINTER(Scene), f32p, f32p]
        L.cvro_camera_ray.argtypes = [C.POINTER(Camera), C.c_uint32, C.c_float, C.c_float, f32p, f32p]
        L.cvro_utilhash.argtypes = [C.c_uint32]
        L.cvro_utilhash.restype = C.c_uint32
        L.cvro_morton3d.argtypes = [C.c_float] * 3
        L.cvro_morton3d.restype = C.c_uint32
        L.cvro_tile_table.argtypes = [C.c_uint32] * 4 + [u32p, u32p]
        L.cvro_default_camera.argtypes = [C.c_uint32, C.c_uint32, C.c_float, f32p, f32p]
        _lib = L
    return _lib


_ref_host = None


def ref_host():
    """The reference's own host-compiled functions, or None when not built."""
    global _ref_host
    if _ref_host is None:
        if not os.path.exists(REF_HOST_SO):
            return None
        R = C.CDLL(REF_HOST_SO)
        R.ref_aabb_intersect.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.POINTER(C.c_int)]
        R.ref_aabb_intersect.restype = C.c_int
        R.ref_aabb_transform.argtypes = [f32p, f32p, f32p]
        R.ref_frame_from_z.argtypes = [f32p] * 4
        R.ref_frame_local_world.argtypes = [f32p] * 4
        R.ref_hg_sample.argtypes = [f32p, C.c_float, C.c_float, C.c_float, f32p]
        R.ref_ggx_sample.argtypes = [f32p, C.c_float, f32p, f32p, f32p, f32p, C.POINTER(C.c_int)]
        R.ref_ggx_sample.restype = C.c_int
        R.ref_ggx_defaults.argtypes = [f32p, f32p]
        R.ref_fresnel_dielectric.argtypes = [C.c_float, C.c_float, f32p]
        R.ref_fresnel_dielectric.restype = C.c_float
        R.ref_ggx_g1.argtypes = [f32p, f32p, f32p]
        R.ref_ggx_g1.restype = C.c_float
        R.ref_morton3d.argtypes = [C.c_float] * 3
        R.ref_morton3d.restype = C.c_uint32
        R.ref_scale.argtypes = [C.c_float, C.c_float]
        R.ref_scale.restype = C.c_float
        R.ref_fmaxf3.argtypes = [C.c_float] * 3
        R.ref_fmaxf3.restype = C.c_float
        _ref_host = R
    return _ref_host


_ref_cpu = None


def ref_cpu():
    """The reference's own naiveSK / regenerationSK kernels compiled for the host by g++
    (oracle/ref_cpu_harness.cpp), or None when not built."""
    global _ref_cpu
    if _ref_cpu is None:
        if not os.path.exists(REF_CPU_SO):
            return None
        R = C.CDLL(REF_CPU_SO)
        R.refcpu_set_scene.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int,
                                       f32p, f32p, C.c_float, C.c_float]
        R.refcpu_set_scene.restype = None
        R.refcpu_set_camera.argtypes = [f32p, f32p, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_uint, C.c_uint]
        R.refcpu_set_camera.restype = None
        R.refcpu_ggx_defaults.argtypes = [f32p, f32p]
        R.refcpu_render_naive.argtypes = [C.c_uint64, C.c_uint64, f32p, C.c_int]
        R.refcpu_render_naive.restype = None
        R.refcpu_trace_paths_naive.argtypes = [C.c_uint64, C.c_uint64, f32p, C.c_int]
        R.refcpu_trace_paths_naive.restype = None
        R.refcpu_render_regen.argtypes = [C.c_uint64, C.c_uint, f32p, C.c_int]
        R.refcpu_render_regen.restype = None
        R.refcpu_density.argtypes = [f32p]
        R.refcpu_density.restype = C.c_float
        R.refcpu_albedo_at_world.argtypes = [f32p, f32p]
        R.refcpu_camera_ray.argtypes = [C.c_uint, C.c_int, f32p, f32p]
        R.refcpu_woodcock.argtypes = [f32p, f32p, C.c_float, C.c_int, C.POINTER(C.c_int)]
        R.refcpu_woodcock.restype = C.c_float
        R.refcpu_trace_one_naive.argtypes = [C.c_uint, f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        R.refcpu_trace_one_naive.restype = None
        R.refcpu_ggx_g1.argtypes = [f32p] * 3
        R.refcpu_ggx_g1.restype = C.c_float
        R.refcpu_ggx_sample.argtypes = [f32p, C.c_int, f32p, f32p]
        R.refcpu_ggx_sample.restype = C.c_int
        _ref_cpu = R
    return _ref_cpu


class RefCpu:
    """One scene + camera loaded into the host-compiled reference kernels (process-global
    state, like the reference's __constant__ symbols: one instance at a time)."""

    def __init__(self, scene: "Scene", cam: "Camera"):
        R = ref_cpu()
        if R is None:
            raise RuntimeError("oracle/_ref/libcvr_ref_cpu.so is not built (needs /root/reference)")
        self.R = R
        self._keep = scene._keep  # the arrays are borrowed by the library
        R.refcpu_set_scene(scene.density, scene.dnx, scene.dny, scene.dnz, scene.albedo, scene.anx, scene.any,
                           scene.anz, scene.box_min, scene.box_max, scene.scale, scene.max_density)
        self.w, self.h = int(cam.resolution[0]), int(cam.resolution[1])
        R.refcpu_set_camera(cam.inv_view, cam.raster_to_view, self.w, self.h, cam.pixel_index_range[0],
                            cam.pixel_index_range[1], cam.offset[0], cam.offset[1])

    def trace_paths_naive(self, first: int, count: int, n_threads: int | None = None) -> np.ndarray:
        out = np.zeros((count, 4), np.float32)
        self.R.refcpu_trace_paths_naive(first, count, fp(out), n_threads or os.cpu_count() or 1)
        return out

    def trace_one_naive(self, tid: int):
        """(rgba, uniforms drawn, texels fetched) of one naiveSK path."""
        out = np.zeros(4, np.float32)
        nd, nt = C.c_uint64(), C.c_uint64()
        self.R.refcpu_trace_one_naive(tid, fp(out), C.byref(nd), C.byref(nt))
        return out, int(nd.value), int(nt.value)

    def render_naive(self, iterations: int, n_threads: int | None = None) -> np.ndarray:
        out = np.zeros((self.h, self.w, 4), np.float32)
        self.R.refcpu_render_naive(0, self.w * self.h * iterations, fp(out), n_threads or os.cpu_count() or 1)
        return out

    def render_regen(self, iterations: int, seed: int = 0, n_threads: int | None = None) -> np.ndarray:
        out = np.zeros((self.h, self.w, 4), np.float32)
        self.R.refcpu_render_regen(self.w * self.h * iterations, seed, fp(out), n_threads or os.cpu_count() or 1)
        return out


def fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(f32p)


def make_scene(density: np.ndarray, albedo: np.ndarray, box_min, box_max, scale: float,
               max_density: float, hg_g: float = 0.0, ggx_alpha=(0.1, 0.1),
               ggx_eta: float | None = None) -> Scene:
    """density: (nz,ny,nx) float32; albedo: (nz,ny,nx,4) float32 (x fastest in memory)."""
    assert density.dtype == np.float32 and density.ndim == 3 and density.flags["C_CONTIGUOUS"]
    assert albedo.dtype == np.float32 and albedo.ndim == 4 and albedo.shape[3] == 4
    assert albedo.flags["C_CONTIGUOUS"]
    s = Scene()
    s._keep = (density, albedo)
    s.density = fp(density)
    s.dnz, s.dny, s.dnx = density.shape
    s.albedo = fp(albedo)
    s.anz, s.any, s.anx = albedo.shape[:3]
    s.box_min[:] = [float(x) for x in box_min]
    s.box_max[:] = [float(x) for x in box_max]
    s.scale = scale
    s.max_density = max_density
    s.hg_g = hg_g
    s.ggx_alpha[:] = ggx_alpha
    if ggx_eta is None:  # Bsdf.h:21-22: float(1.05) / float(1.01)
        ggx_eta = float(np.float32(1.05) / np.float32(1.01))
    s.ggx_eta = ggx_eta
    return s


def make_camera(tile_w: int, tile_h: int, full_w: int, full_h: int, off_x: int = 0, off_y: int = 0,
                fov_x: float = 0.7, inv_view=None, raster_to_view=None) -> Camera:
    cam = Camera()
    iv = np.zeros(12, np.float32)
    rtv = np.zeros(2, np.float32)
    lib().cvro_default_camera(full_w, full_h, fov_x, fp(iv), fp(rtv))
    if inv_view is not None:
        iv = np.asarray(inv_view, np.float32).reshape(12)
    if raster_to_view is not None:
        rtv = np.asarray(raster_to_view, np.float32).reshape(2)
    cam.inv_view[:] = iv.tolist()
    cam.raster_to_view[:] = rtv.tolist()
    cam.resolution[:] = [float(tile_w), float(tile_h)]
    cam.pixel_index_range[:] = [float(full_w), float(full_h)]
    cam.offset[:] = [off_x, off_y]
    return cam


def xorwow_words(seed: int, n: int) -> np.ndarray:
    r = Rng()
    L = lib()
    L.cvro_rng_init(C.byref(r), C.c_int32(seed))
    return np.array([L.cvro_rng_u32(C.byref(r)) for _ in range(n)], dtype=np.uint32)


def xorwow_floats(seed: int, n: int) -> np.ndarray:
    r = Rng()
    L = lib()
    L.cvro_rng_init(C.byref(r), C.c_int32(seed))
    return np.array([L.cvro_rng_float(C.byref(r)) for _ in range(n)], dtype=np.float32)


def render_naive(scene: Scene, cam: Camera, iterations: int, n_threads: int | None = None):
    w, h = int(cam.resolution[0]), int(cam.resolution[1])
    out = np.zeros((h, w, 4), np.float32)
    ctr = Counters()
    lib().cvro_render_naive(C.byref(scene), C.byref(cam), iterations, fp(out),
                            n_threads or os.cpu_count() or 1, C.byref(ctr))
    return out, ctr.as_dict()


def trace_paths_naive(scene: Scene, cam: Camera, first: int, count: int, n_threads: int | None = None):
    out = np.zeros((count, 4), np.float32)
    ctr = Counters()
    lib().cvro_trace_paths_naive(C.byref(scene), C.byref(cam), first, count, fp(out),
                                 n_threads or os.cpu_count() or 1, C.byref(ctr))
    return out, ctr.as_dict()


def trace_paths_seeded(scene: Scene, cam: Camera, first: int, count: int, seed: int, variant: int = 0,
                       n_threads: int | None = None):
    """Per-path radiances with Rng(seed + path id); variant 0 = naive (scatter pull-back), 1 = regeneration."""
    out = np.zeros((count, 4), np.float32)
    ctr = Counters()
    lib().cvro_trace_paths_seeded(C.byref(scene), C.byref(cam), first, count, seed & 0xffffffff, variant, fp(out),
                                  n_threads or os.cpu_count() or 1, C.byref(ctr))
    return out, ctr.as_dict()


def trace_path_logged(scene: Scene, cam: Camera, rng_seed: int, image_id: int, variant: int = 0,
                      cap_events: int = 4096, cap_dec: int = 1 << 18) -> dict:
    """One path with its event and decision log (cvro_trace_path_logged).  rng_seed is the
    int32 the generator is seeded with (path id + stream base, wrapped)."""
    ev_code, ev_d = np.zeros(cap_events, np.uint32), np.zeros(cap_events, np.uint32)
    dk, dd = np.zeros(cap_dec, np.uint32), np.zeros(cap_dec, np.uint32)
    da, db = np.zeros(cap_dec, np.float32), np.zeros(cap_dec, np.float32)
    t = Trace()
    t.cap_events, t.cap_dec = cap_events, cap_dec
    t.ev_code, t.ev_d = ev_code.ctypes.data_as(u32p), ev_d.ctypes.data_as(u32p)
    t.dec_kind, t.dec_d = dk.ctypes.data_as(u32p), dd.ctypes.data_as(u32p)
    t.dec_a, t.dec_b = fp(da), fp(db)
    rad = np.zeros(3, np.float32)
    seed32 = ((int(rng_seed) + 2 ** 31) % 2 ** 32) - 2 ** 31
    esc = lib().cvro_trace_path_logged(C.byref(scene), C.byref(cam), seed32, image_id, variant, 0, fp(rad), C.byref(t))
    ne, nd = min(t.n_events, cap_events), min(t.n_dec, cap_dec)
    return {"escaped": bool(esc), "radiance": rad, "d0": int(t.d0), "n_events": int(t.n_events), "n_dec": int(t.n_dec),
            "ev_code": ev_code[:ne].copy(), "ev_d": ev_d[:ne].copy(), "dec_kind": dk[:nd].copy(), "dec_d": dd[:nd].copy(),
            "dec_a": da[:nd].copy(), "dec_b": db[:nd].copy()}


def render_regen(scene: Scene, cam: Camera, iterations: int, seed: int = 0, rng_mode: int = 1,
                 n_persistent: int = 4096, n_threads: int | None = None):
    w, h = int(cam.resolution[0]), int(cam.resolution[1])
    out = np.zeros((h, w, 4), np.float32)
    ctr = Counters()
    lib().cvro_render_regen(C.byref(scene), C.byref(cam), iterations, seed, rng_mode, n_persistent,
                            fp(out), n_threads or os.cpu_count() or 1, C.byref(ctr))
    return out, ctr.as_dict()


def tile_table(res_x: int, res_y: int, ntx: int, nty: int):
    dim = np.zeros(2, np.uint32)
    org = np.zeros((ntx * nty, 2), np.uint32)
    lib().cvro_tile_table(res_x, res_y, ntx, nty, dim.ctypes.data_as(u32p), org.ctypes.data_as(u32p))
    return dim, org
