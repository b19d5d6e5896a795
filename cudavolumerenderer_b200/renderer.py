"""Python mirror of CudaVolPath<Launcher> (reference implementation/src/CudaVolPath.{h,cpp})
and TilingConfig (Config.h:61-78): the tile scheduler and progressive-render interface
(AbstractRenderer.h:8-24) driving a launcher through the C ABI call by call, in the
reference's order.  Device buffers are torch CUDA tensors (plumbing only); the
one-call equivalent is VolPTKernelLauncher.renderImage / cvr_render_image.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .launcher import Scene, VolPTKernelLauncher, createLauncher


class TilingConfig:
    """Config.h:61-78.  tile_dim floors (integer division inside ceil, Q6)."""

    def __init__(self, resolution=(400, 400), n_tiles=(1, 1)):
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.n_tiles = (int(n_tiles[0]), int(n_tiles[1]))
        dim, org = abi.tile_table(self.resolution[0], self.resolution[1], self.n_tiles[0], self.n_tiles[1])
        self.tile_dim = (int(dim[0]), int(dim[1]))
        self.tiles = [(int(x), int(y)) for x, y in org]  # CudaVolPath::initTileArray (CudaVolPath.cpp:12-29)


class CudaVolPath:
    """Progressive renderer: initRendering / runIterations / getImage / imageComplete /
    setNIterations / render, as CudaVolPath.cpp:240-295,338-347."""

    def __init__(self, scene: Scene, kernel: str = "regenerationSK", resolution=(1024, 1024),
                 n_tiles=(1, 1), iterations: int = 20, device: int = 0, launcher: VolPTKernelLauncher | None = None,
                 **options):
        import torch

        self._torch = torch
        self.scene = scene
        self.tiling = TilingConfig(resolution, n_tiles)
        self.iterations = int(iterations)
        self.device = device
        self.kernel_launcher = launcher or createLauncher(kernel, device, **options)
        kl = self.kernel_launcher
        tw, th = self.tiling.tile_dim
        # constructor order of CudaVolPath.cpp:39-58
        self.inv_view, rtv = abi.default_camera(self.tiling.resolution[0], self.tiling.resolution[1], scene.fov_x)
        if scene.inv_view is not None:
            self.inv_view = scene.inv_view
        kl.copyRasterToView(float(rtv[0]), float(rtv[1]))
        kl.setResolution(tw, th)
        kl.copyPixelIndexRange(float(self.tiling.resolution[0]), float(self.tiling.resolution[1]))
        kl.init()
        # allocateDeviceMemory (CudaVolPath.cpp:211-232): tile-sized float4 buffer, caller-owned
        self.d_output = torch.zeros((th, tw, 4), dtype=torch.float32, device=f"cuda:{device}")
        self.d_image = torch.zeros((self.tiling.resolution[1], self.tiling.resolution[0], 4),
                                   dtype=torch.float32, device=f"cuda:{device}")
        kl.setStream(torch.cuda.current_stream(device).cuda_stream)
        kl.setOutputPtr(self.d_output.data_ptr())
        kl.allocateDeviceMemory()
        kl.setScene(scene)  # initDeviceScene
        self.current_iteration = 0
        self._tile = 0

    # AbstractProgressiveRenderer ------------------------------------------------
    def setNIterations(self, n: int) -> None:
        self.iterations = int(n)
        self.kernel_launcher.setNIterations(self.iterations)

    def initRendering(self) -> None:
        self.kernel_launcher.copyInvViewMatrix(self.inv_view)  # initCamera (CudaVolPath.cpp:66-85)
        self.current_iteration = 0                              # initRenderState (:202-208)
        self.d_output.zero_()
        self._tile = 0

    def imageComplete(self) -> bool:
        return self._tile == len(self.tiling.tiles)

    def runIterations(self) -> None:
        if self._tile == len(self.tiling.tiles):
            self._tile = 0
        if self._tile == 0:
            self.current_iteration += self.kernel_launcher.getNIterations()
        ox, oy = self.tiling.tiles[self._tile]
        self.kernel_launcher.copyOffset(ox, oy)
        self.kernel_launcher.launchRender()
        self._tile += 1

    def getImage(self, host_image: np.ndarray | None = None) -> np.ndarray:
        """Resolve (divide by current_iteration) the tile just rendered into the image
        and prepare for the next tile (CudaVolPath.cpp:282-295, 188-200)."""
        tw, th = self.tiling.tile_dim
        W, H = self.tiling.resolution
        ox, oy = self.tiling.tiles[self._tile - 1]
        kl = self.kernel_launcher
        kl.resolveTile(self.d_output.data_ptr(), tw, th, self.d_image.data_ptr(), W, H, ox, oy,
                       float(self.current_iteration))
        kl.reset()  # sync + seed advance
        if len(self.tiling.tiles) != 1:  # with one tile the buffer keeps accumulating
            self.d_output.zero_()
        if host_image is None:
            host_image = np.zeros((H, W, 4), np.float32)
        tile = self.d_image[oy:oy + th, ox:ox + tw].cpu().numpy()
        host_image[oy:oy + th, ox:ox + tw] = tile
        return host_image

    def render(self, host_image: np.ndarray | None = None) -> np.ndarray:
        W, H = self.tiling.resolution
        if host_image is None:
            host_image = np.zeros((H, W, 4), np.float32)
        self.setNIterations(self.iterations)
        self.initRendering()
        while not self.imageComplete():
            self.runIterations()
            self.getImage(host_image)
        return host_image

    def close(self):
        self.kernel_launcher.close()
