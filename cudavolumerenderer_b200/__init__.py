"""cudavolumerenderer_b200 -- B200-native replacement for the CudaVolumeRenderer
path-tracing hot path.  The product is libcvr_b200.so (C ABI, include/cvr_abi.h);
this package is the thin host-side mirror of the reference's launcher / renderer
interface over that ABI.  Nothing here computes on the CPU."""
from .abi import CvrError, load  # noqa: F401
from .launcher import (KERNELS, DeviceGroup, NaiveVolPTsk, ProceduralScene, RegenerationVolPTsk, Scene, SortingVolPTsk,  # noqa: F401
                       SparseScene, StreamingVolPTmk, StreamingVolPTsk, VolPTKernelLauncher, createLauncher)
from .renderer import CudaVolPath, TilingConfig  # noqa: F401
from . import scenes  # noqa: F401
