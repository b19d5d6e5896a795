"""CPU-only checks of the C-ABI library: it loads, exports every symbol that
include/cvr_abi.h declares, its GPU-free helpers agree with the oracle, and GPU entry
points fail LOUDLY (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "cudavolumerenderer_b200", "libcvr_b200.so")):
        g.build()
    from cudavolumerenderer_b200 import abi

    return abi.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from cudavolumerenderer_b200 import abi

    header = open(os.path.join(ROOT, "include", "cvr_abi.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cvr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 60
    bound = {name for name, _, _ in abi.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cvr_abi_version() == 3


def test_no_oracle_in_product_path():
    """The product package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "cudavolumerenderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the CPU oracle", ""), os.path.join(dirpath, f)
                assert "cvro_" not in src, os.path.join(dirpath, f)


def test_tile_table_and_camera_match_oracle(lib, oracle):
    from cudavolumerenderer_b200 import abi

    for (rx, ry, tx, ty) in [(1024, 1024, 10, 10), (256, 256, 1, 1), (4096, 4096, 7, 3), (50, 50, 3, 3),
                             (1920, 1080, 64, 64)]:
        d1, o1 = abi.tile_table(rx, ry, tx, ty)
        d2, o2 = oracle.tile_table(rx, ry, tx, ty)
        assert np.array_equal(d1, d2) and np.array_equal(o1, o2)
    for (rx, ry, fov) in [(1024, 1024, 0.7), (400, 300, 0.33), (4096, 2048, 45.0)]:
        iv, rtv = abi.default_camera(rx, ry, fov)
        cam = oracle.make_camera(rx, ry, rx, ry, fov_x=fov)
        assert np.array_equal(iv, np.array(cam.inv_view[:], np.float32))
        assert np.array_equal(rtv, np.array(cam.raster_to_view[:], np.float32))


def test_synth_scenes_are_deterministic_and_shaped(lib):
    from cudavolumerenderer_b200 import scenes

    a, b = scenes.bucky(), scenes.bucky()
    assert a.density.shape == (32, 32, 32) and np.array_equal(a.density, b.density)
    assert a.max_density == 1.0 and a.scale == 40.0 and a.density.max() == 1.0
    # Raw loader quantisation: density = byte / max byte
    assert len(np.unique(a.density)) <= 256
    assert a.albedo.shape == (32, 32, 32, 4) and np.all(a.albedo[..., 3] == 1.0)
    # empty voxels take transfer-function entry 0 = (0.02, 0.2, 0.02)
    assert np.allclose(a.albedo[a.density == 0][0, :3], [0.02, 0.2, 0.02])
    h = scenes.hetvol()
    assert h.density.shape == (50, 128, 128) and h.box_max == pytest.approx((0.64, 0.64, 0.25))
    assert 0 < h.max_density <= 1.0 and h.scale == 800.0 and h.fov_x == pytest.approx(0.33)
    assert np.allclose(h.albedo[0, 0, 0], [0.96, 0.84, 0.68, 1.0])
    m = scenes.manix(dims=(32, 28, 32))
    assert m.density.shape == (32, 28, 32) and np.array_equal(m.albedo[..., 0], m.density)
    assert np.all(m.albedo[..., 1:3] == 0)


def test_gpu_entry_points_fail_loudly_without_a_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.cvr_create(b"regenerationSK", 0, C.byref(h)) != 0
    msg = lib.cvr_last_error(None).decode()
    assert "no CPU fallback" in msg
    from cudavolumerenderer_b200 import CvrError, RegenerationVolPTsk

    with pytest.raises(CvrError):
        RegenerationVolPTsk(0)


def test_kernel_names_follow_reference_config(lib):
    from cudavolumerenderer_b200 import KERNELS, createLauncher

    # Config::getKernelNamesOrderedVector (Config.h:210-213) minus naiveMK (a different estimator variant)
    assert set(KERNELS) == {"naiveSK", "regenerationSK", "streamingSK", "streamingMK", "sortingSK"}
    with pytest.raises(ValueError):
        createLauncher("cpuSK")
