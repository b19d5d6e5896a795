"""Python mirror of the reference's kernel-launcher plugin surface, over the C ABI.

Class and method names follow RenderKernelLauncher / VolPTKernelLauncher
(reference implementation/src/RenderKernelLauncher.h:20-73) and the three launcher
classes the hot path covers (:75-82 NaiveVolPTsk, :104-113 RegenerationVolPTsk,
:139-151 StreamingVolPTsk) so that code written against the reference reads the same.
Every method is one C-ABI call; errors raise CvrError (the reference exit()s).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .abi import CvrError, Counters, SceneDesc, SparseDesc


class Scene:
    """Host-side scene = what SceneAssembler::getScene returns (Scene.h:19-54): a
    camera (fov / optional matrices) and a HostMedium (Medium.h:109-116,191)."""

    def __init__(self, density: np.ndarray, albedo: np.ndarray | None, box_min, box_max, scale: float,
                 max_density: float, fov_x: float = 0.7, albedo_const=(1.0, 1.0, 1.0), hg_g: float = 0.0,
                 ggx_alpha=(0.1, 0.1), ggx_eta: float | None = None, inv_view=None, name: str = "scene"):
        density = np.ascontiguousarray(density, np.float32)
        if density.ndim != 3:
            raise ValueError("density must be (nz, ny, nx)")
        if albedo is not None:
            albedo = np.ascontiguousarray(albedo, np.float32)
            if albedo.ndim != 4 or albedo.shape[3] != 4:
                raise ValueError("albedo must be (nz, ny, nx, 4)")
        self.density, self.albedo = density, albedo
        self.box_min = tuple(float(x) for x in box_min)
        self.box_max = tuple(float(x) for x in box_max)
        self.scale, self.max_density = float(scale), float(max_density)
        self.fov_x = float(fov_x)
        self.albedo_const = tuple(float(x) for x in albedo_const)
        self.hg_g = float(hg_g)
        self.ggx_alpha = tuple(float(x) for x in ggx_alpha)
        # Bsdf.h:21-22: float(1.05) / float(1.01)
        self.ggx_eta = float(np.float32(1.05) / np.float32(1.01)) if ggx_eta is None else float(ggx_eta)
        self.inv_view = None if inv_view is None else np.asarray(inv_view, np.float32).reshape(12)
        self.name = name

    def desc(self) -> SceneDesc:
        d = SceneDesc()
        d.density = self.density.ctypes.data
        d.density_dim[:] = [self.density.shape[2], self.density.shape[1], self.density.shape[0]]
        if self.albedo is not None:
            d.albedo = self.albedo.ctypes.data
            d.albedo_dim[:] = [self.albedo.shape[2], self.albedo.shape[1], self.albedo.shape[0]]
        else:
            d.albedo = None
        d.albedo_const[:] = self.albedo_const
        d.box_min[:] = self.box_min
        d.box_max[:] = self.box_max
        d.scale, d.max_density, d.hg_g = self.scale, self.max_density, self.hg_g
        d.ggx_alpha[:] = self.ggx_alpha
        d.ggx_eta = self.ggx_eta
        d.density_on_device = 0
        return d

    @property
    def volume_bytes(self) -> int:
        return self.density.nbytes + (self.albedo.nbytes if self.albedo is not None else 0)


class SparseScene:
    """A VDB-style sparse medium: 8^3 leaves re-laid into device bricks without densifying
    (cvr_set_scene_sparse).  Same medium scalars as Scene; constant albedo."""

    def __init__(self, origins: np.ndarray, values: np.ndarray, dim, bbox_min, box_min=(-0.5,) * 3, box_max=(0.5,) * 3,
                 scale: float = 100.0, max_density: float = 0.0, fov_x: float = 0.7, albedo_const=(1.0, 1.0, 1.0),
                 hg_g: float = 0.0, name: str = "sparse"):
        self.origins = np.ascontiguousarray(origins, np.int32).reshape(-1, 3)
        self.values = np.ascontiguousarray(values, np.float32).reshape(-1, 512)
        if len(self.origins) != len(self.values):
            raise ValueError("one origin per leaf")
        self.dim = tuple(int(v) for v in dim)
        self.bbox_min = tuple(int(v) for v in bbox_min)
        self.box_min, self.box_max = tuple(map(float, box_min)), tuple(map(float, box_max))
        self.scale, self.max_density, self.fov_x = float(scale), float(max_density), float(fov_x)
        self.albedo_const = tuple(map(float, albedo_const))
        self.hg_g = float(hg_g)
        self.inv_view = None
        self.name = name

    def desc(self) -> SparseDesc:
        d = SparseDesc()
        d.dim[:] = self.dim
        d.bbox_min[:] = self.bbox_min
        d.n_leaves = len(self.origins)
        d.leaf_origins = self.origins.ctypes.data
        d.leaf_values = self.values.ctypes.data
        d.albedo_const[:] = self.albedo_const
        d.box_min[:] = self.box_min
        d.box_max[:] = self.box_max
        d.scale, d.max_density, d.hg_g = self.scale, self.max_density, self.hg_g
        d.ggx_alpha[:] = (0.1, 0.1)
        d.ggx_eta = float(np.float32(1.05) / np.float32(1.01))
        return d


class ProceduralScene:
    """A volume generated on the device (cvr_set_scene_procedural): kind "fbm" (dense) or
    "sparsefbm" (VDB-style, ~3 % of the bricks active), n^3 index space."""

    def __init__(self, kind: str, n: int, seed: int = 0, box_min=(-0.5,) * 3, box_max=(0.5,) * 3, scale: float = 100.0,
                 max_density: float = 0.0, fov_x: float = 0.7, albedo_const=(0.99,) * 3, hg_g: float = 0.0):
        self.kind, self.n, self.seed = kind, int(n), int(seed)
        self.box_min, self.box_max = tuple(map(float, box_min)), tuple(map(float, box_max))
        self.scale, self.max_density, self.fov_x = float(scale), float(max_density), float(fov_x)
        self.albedo_const = tuple(map(float, albedo_const))
        self.hg_g = float(hg_g)
        self.inv_view = None
        self.name = f"{kind}{n}"

    def desc(self) -> SceneDesc:
        d = SceneDesc()
        d.albedo_const[:] = self.albedo_const
        d.box_min[:] = self.box_min
        d.box_max[:] = self.box_max
        d.scale, d.max_density, d.hg_g = self.scale, self.max_density, self.hg_g
        d.ggx_alpha[:] = (0.1, 0.1)
        d.ggx_eta = float(np.float32(1.05) / np.float32(1.01))
        return d


class VolPTKernelLauncher:
    """RenderKernelLauncher + VolPTKernelLauncher<DeviceScene> (RenderKernelLauncher.h:20-73)."""

    KERNEL = "regenerationSK"

    def __init__(self, device: int = 0, **options):
        self._lib = abi.load()
        h = C.c_void_p()
        rc = self._lib.cvr_create(self.KERNEL.encode(), device, C.byref(h))
        if rc:
            raise CvrError("cvr_create: " + self._lib.cvr_last_error(None).decode())
        self._h = h
        self.device = device
        for k, v in options.items():
            self.setOption(k, v)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        abi.check(self._h, rc, what)

    # -- options / stream (new-build additions)
    def setOption(self, key: str, value) -> None:
        self._ck(self._lib.cvr_set_option(self._h, key.encode(), str(value).encode()), f"set_option({key})")

    def getOption(self, key: str) -> str:
        buf = C.create_string_buffer(64)
        self._ck(self._lib.cvr_get_option(self._h, key.encode(), buf, 64), f"get_option({key})")
        return buf.value.decode()

    def setStream(self, cuda_stream_ptr: int | None) -> None:
        """cudaStream_t to launch on.  None = the handle's own (non-blocking) stream.  0 -- what
        torch reports for its DEFAULT stream -- is passed on as cudaStreamLegacy (0x1): the C ABI
        reads a NULL stream as "own stream", and work launched there would not be ordered with
        torch operations on the default stream (a d_image.zero_() enqueued before the render could
        execute after the tiles were resolved and wipe them)."""
        if cuda_stream_ptr is not None and int(cuda_stream_ptr) == 0:
            cuda_stream_ptr = 1  # cudaStreamLegacy
        self._stream = cuda_stream_ptr
        self._ck(self._lib.cvr_set_stream(self._h, cuda_stream_ptr), "set_stream")

    def streamPtr(self) -> int | None:
        """The stream handle given to setStream (cudaStreamLegacy = 1 for the default stream), or
        None while the handle launches on its own stream."""
        return getattr(self, "_stream", None)

    # -- RenderKernelLauncher.h:31-51
    def setOutputPtr(self, d_output: int) -> None:
        self._ck(self._lib.cvr_set_output(self._h, d_output), "setOutputPtr")

    def setResolution(self, w: int, h: int) -> None:
        self._ck(self._lib.cvr_set_resolution(self._h, w, h), "setResolution")

    def copyInvViewMatrix(self, m) -> None:
        m = np.ascontiguousarray(m, np.float32).reshape(12)
        self._ck(self._lib.cvr_set_inv_view_matrix(self._h, m.ctypes.data_as(abi.f32p)), "copyInvViewMatrix")

    def copyRasterToView(self, x: float, y: float) -> None:
        self._ck(self._lib.cvr_set_raster_to_view(self._h, x, y), "copyRasterToView")

    def copyPixelIndexRange(self, w: float, h: float) -> None:
        self._ck(self._lib.cvr_set_pixel_index_range(self._h, w, h), "copyPixelIndexRange")

    def copyOffset(self, x: int, y: int) -> None:
        self._ck(self._lib.cvr_set_offset(self._h, x, y), "copyOffset")

    def init(self) -> None:
        self._ck(self._lib.cvr_init(self._h), "init")

    def allocateDeviceMemory(self) -> None:
        self._ck(self._lib.cvr_allocate(self._h), "allocateDeviceMemory")

    def launchRender(self) -> None:
        self._ck(self._lib.cvr_launch_render(self._h), "launchRender")

    def reset(self) -> None:
        self._ck(self._lib.cvr_reset(self._h), "reset")

    def sync(self) -> None:
        self._ck(self._lib.cvr_sync(self._h), "sync")

    def releaseDeviceMemory(self) -> None:
        self._ck(self._lib.cvr_release(self._h), "releaseDeviceMemory")

    # -- VolPTKernelLauncher (RenderKernelLauncher.h:54-73)
    def setNIterations(self, n: int) -> None:
        self._ck(self._lib.cvr_set_iterations(self._h, n), "setNIterations")

    def getNIterations(self) -> int:
        n = C.c_uint32()
        self._ck(self._lib.cvr_get_iterations(self._h, C.byref(n)), "getNIterations")
        return n.value

    def setScene(self, scene) -> None:
        d = scene.desc()
        if isinstance(scene, SparseScene):
            self._ck(self._lib.cvr_set_scene_sparse(self._h, C.byref(d)), "setScene(sparse)")
        elif isinstance(scene, ProceduralScene):
            mx = C.c_float()
            self._ck(self._lib.cvr_set_scene_procedural(self._h, scene.kind.encode(), scene.n, scene.seed, C.byref(d),
                                                        C.byref(mx)), "setScene(procedural)")
            scene.max_density = float(mx.value)
        else:
            self._ck(self._lib.cvr_set_scene(self._h, C.byref(d)), "setScene")
        self._scene = scene

    def volumeInfo(self) -> dict:
        b, n, lay = C.c_uint64(), C.c_uint64(), C.c_int32()
        self._ck(self._lib.cvr_get_volume_info(self._h, C.byref(b), C.byref(n), C.byref(lay)), "volumeInfo")
        return {"layout_bytes": b.value, "bricks": n.value, "layout": ("linear", "cell8", "brick")[lay.value]}

    def getScene(self) -> Scene:
        return self._scene

    # -- seeds / sharding / statistics
    def setSeed(self, seed: int) -> None:
        self._ck(self._lib.cvr_set_seed(self._h, seed & 0xffffffff), "setSeed")

    def getSeed(self) -> int:
        s = C.c_uint32()
        self._ck(self._lib.cvr_get_seed(self._h, C.byref(s)), "getSeed")
        return s.value

    def setSampleRange(self, first: int, count: int) -> None:
        self._ck(self._lib.cvr_set_sample_range(self._h, first, count), "setSampleRange")

    def counters(self) -> dict:
        c = Counters()
        self._ck(self._lib.cvr_get_counters(self._h, C.byref(c)), "counters")
        return c.as_dict()

    def resetCounters(self) -> None:
        self._ck(self._lib.cvr_reset_counters(self._h), "resetCounters")

    def launchShape(self):
        g, b, r = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.cvr_get_launch_shape(self._h, C.byref(g), C.byref(b), C.byref(r)), "launchShape")
        return g.value, b.value, r.value

    def resolveTile(self, d_tile: int, tile_w: int, tile_h: int, d_image: int, full_w: int, full_h: int,
                    off_x: int, off_y: int, scale: float) -> None:
        self._ck(self._lib.cvr_resolve_tile(self._h, d_tile, tile_w, tile_h, d_image, full_w, full_h,
                                            off_x, off_y, scale), "resolveTile")

    def resolveTileDisplay(self, d_tile: int, tile_w: int, tile_h: int, d_transfer: int, d_display: int, full_w: int,
                           full_h: int, off_x: int, off_y: int, scale: float, reset_transfer: bool = False) -> None:
        """DeviceTiledImageBufferTansferDelegate::transfer (ImageBufferTransfer.cu:128-157): accumulate the tile
        into the float4 transfer buffer and write gamma-corrected uchar4 display pixels."""
        self._ck(self._lib.cvr_resolve_tile_display(self._h, d_tile, tile_w, tile_h, d_transfer, d_display, full_w, full_h,
                                                    off_x, off_y, scale, 1 if reset_transfer else 0), "resolveTileDisplay")

    def renderImage(self, res, n_tiles=(1, 1), iterations: int = 20, fov_x: float = 0.7, inv_view=None,
                    raster_to_view=None, tile_first: int = 0, tile_stride: int = 1, sample_first: int = 0,
                    sample_count: int = 0, fuse_tiles: bool = False, host_image: np.ndarray | None = None,
                    d_image: int | None = None) -> np.ndarray | None:
        """CudaVolPath::render (CudaVolPath.cpp:338-347) in one C-ABI call."""
        r = abi.RenderDesc()
        r.res_x, r.res_y = res
        r.n_tiles_x, r.n_tiles_y = n_tiles
        r.iterations, r.fov_x = iterations, fov_x
        keep = []
        if inv_view is not None:
            iv = np.ascontiguousarray(inv_view, np.float32).reshape(12)
            keep.append(iv)
            r.inv_view = iv.ctypes.data_as(abi.f32p)
        if raster_to_view is not None:
            rv = np.ascontiguousarray(raster_to_view, np.float32).reshape(2)
            keep.append(rv)
            r.raster_to_view = rv.ctypes.data_as(abi.f32p)
        r.tile_first, r.tile_stride = tile_first, tile_stride
        r.sample_first, r.sample_count = sample_first, sample_count
        r.fuse_tiles = 1 if fuse_tiles else 0
        if host_image is None and d_image is None:
            host_image = np.zeros((res[1], res[0], 4), np.float32)
        hp = host_image.ctypes.data if host_image is not None else None
        self._ck(self._lib.cvr_render_image(self._h, C.byref(r), hp, d_image), "renderImage")
        return host_image

    def renderImageSharded(self, res, n_tiles, iterations: int, shard: "abi.Shard", fov_x: float = 0.7, inv_view=None,
                           fuse_tiles: bool = True, host_image: np.ndarray | None = None, d_image: int | None = None):
        """cvr_render_image_sharded: this rank's share (abi.shard_plan) of the image; pixels it does not
        own are ZERO in the result, which is a term of the sum over ranks."""
        r = abi.RenderDesc()
        r.res_x, r.res_y = res
        r.n_tiles_x, r.n_tiles_y = n_tiles
        r.iterations, r.fov_x = iterations, fov_x
        keep = []
        if inv_view is not None:
            iv = np.ascontiguousarray(inv_view, np.float32).reshape(12)
            keep.append(iv)
            r.inv_view = iv.ctypes.data_as(abi.f32p)
        r.fuse_tiles = 1 if fuse_tiles else 0
        if host_image is None and d_image is None:
            host_image = np.zeros((res[1], res[0], 4), np.float32)
        hp = host_image.ctypes.data if host_image is not None else None
        self._ck(self._lib.cvr_render_image_sharded(self._h, C.byref(r), C.byref(shard), hp, d_image), "renderImageSharded")
        return host_image

    # -- parity hooks
    def tracePaths(self, first: int, count: int, d_per_path: int) -> None:
        self._ck(self._lib.cvr_trace_paths(self._h, first, count, d_per_path), "tracePaths")

    def tracePathsLogged(self, first: int, count: int, d_per_path: int, d_log: int, log_cap: int) -> None:
        """tracePaths plus the per-path event log (DEVICE uint2[count * log_cap], include/cvr_abi.h)."""
        self._ck(self._lib.cvr_trace_paths_logged(self._h, first, count, d_per_path, d_log, log_cap), "tracePathsLogged")

    def trigCheck(self, limit: float = 8.0):
        """cvr_debug_trig_check: (mismatch counts of sin / cos / tan, first differing |x| bit patterns)."""
        m = (C.c_uint64 * 3)()
        f = (C.c_uint32 * 3)()
        self._ck(self._lib.cvr_debug_trig_check(self._h, limit, m, f), "trigCheck")
        return [int(v) for v in m], [int(v) for v in f]

    def rngKat(self, seeds, n: int):
        seeds = np.ascontiguousarray(seeds, np.int32)
        w = np.zeros((len(seeds), n), np.uint32)
        u = np.zeros((len(seeds), n), np.float32)
        self._ck(self._lib.cvr_rng_kat(self._h, seeds.ctypes.data_as(C.POINTER(C.c_int32)), len(seeds), n,
                                       w.ctypes.data_as(abi.u32p), u.ctypes.data_as(abi.f32p)), "rngKat")
        return w, u

    def gatherRoofline(self, footprint_bytes: int, loads_per_thread: int = 256, unroll: int = 8, bypass_l1: bool = False) -> float:
        """Measured random 32-byte-sector gather bandwidth (GB/s) at this footprint (bypass_l1: the loads
        are not allocated in the L1, the pure L2 / HBM -> SM sector rate)."""
        g = C.c_double()
        if bypass_l1:
            unroll = -8
        self._ck(self._lib.cvr_gather_roofline(self._h, footprint_bytes, loads_per_thread, unroll, C.byref(g)),
                 "gatherRoofline")
        return g.value

    def debugLookup(self, p01: np.ndarray):
        p = np.ascontiguousarray(p01, np.float32).reshape(-1, 3)
        d = np.zeros(len(p), np.float32)
        a = np.zeros((len(p), 3), np.float32)
        self._ck(self._lib.cvr_debug_lookup(self._h, p.ctypes.data_as(abi.f32p), len(p),
                                            d.ctypes.data_as(abi.f32p), a.ctypes.data_as(abi.f32p)), "debugLookup")
        return d, a


class NaiveVolPTsk(VolPTKernelLauncher):
    """-k naiveSK (RenderKernelLauncher.h:75-82; NaiveVolPTsk_kernel.cuh)."""
    KERNEL = "naiveSK"


class RegenerationVolPTsk(VolPTKernelLauncher):
    """-k regenerationSK (RenderKernelLauncher.h:104-113; RegenerationVolPTsk_kernel.cuh:146-232)."""
    KERNEL = "regenerationSK"


class StreamingVolPTsk(VolPTKernelLauncher):
    """-k streamingSK (RenderKernelLauncher.h:139-151; StreamingVolPTsk_kernel.cuh)."""
    KERNEL = "streamingSK"


class StreamingVolPTmk(VolPTKernelLauncher):
    """-k streamingMK (RenderKernelLauncher.h:115-137; StreamingVolPTmk_kernel.cuh): per-path streams
    Rng(c_seed + path_id) (:55), pull-back at scatter (:194), seed += n_paths per reset
    (RenderKernelLauncher.cu:480-481).  One persistent kernel here, not one launch per bounce."""
    KERNEL = "streamingMK"


class SortingVolPTsk(VolPTKernelLauncher):
    """-k sortingSK (RenderKernelLauncher.h:153-170; SortingVolPTsk_kernel.cuh): the streamingSK
    estimator, streams and seed rule (:227-230, :314; RenderKernelLauncher.cu:664-665); its Morton
    ordering of rays is a scheduling choice, which belongs to this library."""
    KERNEL = "sortingSK"


KERNELS = {"naiveSK": NaiveVolPTsk, "regenerationSK": RegenerationVolPTsk, "streamingSK": StreamingVolPTsk,
           "streamingMK": StreamingVolPTmk, "sortingSK": SortingVolPTsk}


class DeviceGroup:
    """cvr_group_*: one launcher per device of THIS process (one host thread per device inside the
    library), every device holding a replica of the scene, the framebuffers combined with one
    ncclReduce to the first device.  The multi-process form (one process per GPU under torchrun)
    is cudavolumerenderer_b200.distributed.render_sharded."""

    def __init__(self, kernel: str = "regenerationSK", devices=None, n_devices: int | None = None, **options):
        self._lib = abi.load()
        if devices is None:
            devices = list(range(n_devices or 1))
        arr = (C.c_int * len(devices))(*devices)
        g = C.c_void_p()
        if self._lib.cvr_group_create(kernel.encode(), arr, len(devices), C.byref(g)):
            raise CvrError("cvr_group_create: " + self._lib.cvr_group_last_error(None).decode())
        self._g = g
        self.devices = list(devices)
        for k, v in options.items():
            self.setOption(k, v)

    def _ck(self, rc, what):
        if rc:
            raise CvrError(f"{what}: {self._lib.cvr_group_last_error(self._g).decode()}")

    def close(self):
        if getattr(self, "_g", None):
            self._lib.cvr_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setOption(self, key: str, value) -> None:
        self._ck(self._lib.cvr_group_set_option(self._g, key.encode(), str(value).encode()), f"set_option({key})")

    def setSeed(self, seed: int) -> None:
        self._ck(self._lib.cvr_group_set_seed(self._g, seed & 0xffffffff), "setSeed")

    def setScene(self, scene) -> None:
        d = scene.desc()
        if isinstance(scene, SparseScene):
            self._ck(self._lib.cvr_group_set_scene_sparse(self._g, C.byref(d)), "setScene(sparse)")
        elif isinstance(scene, ProceduralScene):
            mx = C.c_float()
            self._ck(self._lib.cvr_group_set_scene_procedural(self._g, scene.kind.encode(), scene.n, scene.seed, C.byref(d),
                                                              C.byref(mx)), "setScene(procedural)")
            scene.max_density = float(mx.value)
        else:
            self._ck(self._lib.cvr_group_set_scene(self._g, C.byref(d)), "setScene")
        self._scene = scene

    def resolveTileDisplay(self, d_tile: int, tile_w: int, tile_h: int, d_transfer: int, d_display: int, full_w: int,
                           full_h: int, off_x: int, off_y: int, scale: float, reset_transfer: bool = False) -> None:
        """DeviceTiledImageBufferTansferDelegate::transfer (ImageBufferTransfer.cu:128-157): accumulate the tile
        into the float4 transfer buffer and write gamma-corrected uchar4 display pixels."""
        self._ck(self._lib.cvr_resolve_tile_display(self._h, d_tile, tile_w, tile_h, d_transfer, d_display, full_w, full_h,
                                                    off_x, off_y, scale, 1 if reset_transfer else 0), "resolveTileDisplay")

    def renderImage(self, res, n_tiles=(1, 1), iterations: int = 20, shard: str = "balanced", fov_x: float = 0.7,
                    fuse_tiles: bool = True, host_image: np.ndarray | None = None) -> np.ndarray:
        r = abi.RenderDesc()
        r.res_x, r.res_y = res
        r.n_tiles_x, r.n_tiles_y = n_tiles
        r.iterations, r.fov_x = iterations, fov_x
        r.fuse_tiles = 1 if fuse_tiles else 0
        if host_image is None:
            host_image = np.zeros((res[1], res[0], 4), np.float32)
        self._ck(self._lib.cvr_group_render_image(self._g, C.byref(r), abi.SHARD_MODES[shard], host_image.ctypes.data, None),
                 "renderImage")
        return host_image

    def counters(self) -> dict:
        c = Counters()
        self._ck(self._lib.cvr_group_get_counters(self._g, C.byref(c)), "counters")
        return c.as_dict()

    def resetCounters(self) -> None:
        self._ck(self._lib.cvr_group_reset_counters(self._g), "resetCounters")


def createLauncher(kernel: str, device: int = 0, **options) -> VolPTKernelLauncher:
    """Config::getKernel + RendererFactory switch (Config.h:228-235, RendererFactory.h:37-115)."""
    if kernel not in KERNELS:
        raise ValueError(f"kernel '{kernel}' not available (choices: {sorted(KERNELS)})")
    return KERNELS[kernel](device, **options)
