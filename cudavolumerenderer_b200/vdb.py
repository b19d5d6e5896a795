"""OpenVDB files through libcvr_b200.so's dependency-free reader (include/cvr_abi.h,
cvr_vdb_*): the Python face of what replaces the reference's VDBAdapter
(implementation/vdb_adapter/VDBAdapter.{h,cpp}).  No OpenVDB, no CPU compute here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class VdbFile:
    """VDBAdapter-shaped view of a .vdb file: loadVDBFile happens in the constructor."""

    def __init__(self, path: str):
        self._lib = abi.load()
        self._h = C.c_void_p()
        if self._lib.cvr_vdb_open(str(path).encode(), C.byref(self._h)):
            raise abi.CvrError(self._lib.cvr_vdb_last_error().decode())

    def close(self) -> None:
        if self._h:
            self._lib.cvr_vdb_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise abi.CvrError(self._lib.cvr_vdb_last_error().decode())

    def grids(self) -> list[dict]:
        n = C.c_int32()
        self._ck(self._lib.cvr_vdb_grid_count(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            inf = abi.VdbGridInfo()
            self._ck(self._lib.cvr_vdb_grid_info(self._h, i, C.byref(inf)))
            out.append({
                "name": inf.name.decode(), "type": inf.type.decode(), "channels": inf.channels,
                "compression": inf.compression, "file_version": inf.file_version,
                "bbox_min": tuple(inf.bbox_min), "bbox_max": tuple(inf.bbox_max), "dim": tuple(inf.dim),
                "background": tuple(inf.background), "active_voxels": inf.active_voxels,
                "leaf_count": inf.leaf_count, "active_tiles": inf.active_tiles,
            })
        return out

    def grid(self, name: str) -> dict:
        for g in self.grids():
            if g["name"] == name:
                return g
        # messages of VDBAdapter::loadVDBFile (VDBAdapter.cpp:21-37)
        raise abi.CvrError(f"VDB file does not contain a{'n' if name[:1] in 'aeiou' else ''} {name} grid")

    def meta(self, grid: str | None, key: str) -> str:
        buf = C.create_string_buffer(1024)
        self._ck(self._lib.cvr_vdb_grid_meta(self._h, grid.encode() if grid else None, key.encode(), buf, 1024))
        return buf.value.decode()

    def getGridResolution(self) -> tuple[int, int, int]:  # VDBAdapter.cpp:46-55
        return self.grid("density")["dim"]

    def densify(self, name: str, out_channels: int | None = None, inactive=None) -> np.ndarray:
        """get{Density,Albedo}DataAsLinearArray (VDBAdapter.cpp:57-114): (nz, ny, nx[, C]) float32."""
        g = self.grid(name)
        ch = out_channels or g["channels"]
        nx, ny, nz = g["dim"]
        out = np.empty((nz, ny, nx, ch), np.float32)
        ina = None
        if inactive is not None:
            ina = (C.c_float * 3)(*[float(v) for v in inactive])
        self._ck(self._lib.cvr_vdb_densify(self._h, name.encode(), ch, ina, out.ctypes.data_as(abi.f32p), out.size))
        return out[..., 0] if ch == 1 else out

    def leaves(self, name: str):
        """The sparse form: (origins (n,3) int32, masks (n,8) uint64, values (n,512[,C]) float32)."""
        g = self.grid(name)
        n, ch = g["leaf_count"], g["channels"]
        org = np.empty((n, 3), np.int32)
        msk = np.empty((n, 8), np.uint64)
        val = np.empty((n, 512, ch), np.float32)
        self._ck(self._lib.cvr_vdb_leaves(self._h, name.encode(), 0, n, org.ctypes.data_as(C.POINTER(C.c_int32)),
                                          msk.ctypes.data_as(C.POINTER(C.c_uint64)), val.ctypes.data_as(abi.f32p)))
        return org, msk, (val[..., 0] if ch == 1 else val)
