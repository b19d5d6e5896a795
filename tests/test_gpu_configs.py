"""BASELINE.json configs at their FULL sizes on the GPU.

C1 (bucky 32^3, 256x256, 16 spp, 1 tile, naive) is small enough for the CPU oracle to
render completely: direct comparison.  C2 (hetvol 1024^2 x 64 spp, regenerationSK) and C3
(MANIX 1024^2 x 256 spp, 10x10 tiles) are checked through size-independent properties
(sharding recomposition, tile/fused equivalence, untouched remainder pixels, exact path
counts) plus an oracle comparison on a bounded sample of the same image."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REAL_VDB = "/root/reference/data/vdb/bonsai_small.vdb"


@pytest.fixture(scope="module")
def cvr():
    import cudavolumerenderer_b200 as pkg

    return pkg


def _oscene(oracle, sc):
    return oracle.make_scene(sc.density, sc.albedo, sc.box_min, sc.box_max, sc.scale, sc.max_density)


def _const_albedo(rgb):
    a = np.empty((2, 2, 2, 4), np.float32)
    a[..., :3] = np.asarray(rgb, np.float32)
    a[..., 3] = 1.0
    return a


def _oracle_sample(cvr, oracle, kl, osc, fov_x, full, tile, off, spp, seed, variant, min_agree):
    """Bounded-sample ORACLE comparison inside a full-size configuration: the paths of one tile of
    the full image (same camera, same streams Rng(seed + path id)) traced on the device through
    the handle's scene as it stands and on the CPU by the oracle; per-path radiances must agree
    within 1e-4 on at least `min_agree` of the paths (ulp-level flips, tests/test_gpu_parity.py)."""
    import torch

    iv, rtv = cvr.abi.default_camera(full[0], full[1], fov_x)
    kl.copyRasterToView(float(rtv[0]), float(rtv[1]))
    kl.setResolution(tile[0], tile[1])
    kl.copyPixelIndexRange(float(full[0]), float(full[1]))
    kl.copyInvViewMatrix(iv)
    kl.copyOffset(*off)
    kl.setNIterations(spp)
    kl.setSampleRange(0, 0)
    kl.setSeed(seed)
    n = tile[0] * tile[1] * spp
    per = torch.zeros((n, 4), dtype=torch.float32, device="cuda:0")
    kl.tracePaths(0, n, per.data_ptr())
    kl.sync()
    got = per.cpu().numpy()
    cam = oracle.make_camera(tile[0], tile[1], full[0], full[1], off_x=off[0], off_y=off[1], fov_x=fov_x)
    ref, _ = oracle.trace_paths_seeded(osc, cam, 0, n, seed, variant)
    same = np.all(np.abs(got[:, :3] - ref[:, :3]) <= 1e-4, axis=1)
    assert same.mean() >= min_agree, same.mean()
    assert abs(float(got[:, :3].mean()) - float(ref[:, :3].mean())) <= 0.02 * float(ref[:, :3].mean()) + 2e-3
    assert int((ref[:, 3] == 1).sum()) > 0
    return float(same.mean())


def test_c1_bucky_full_config_against_cpu_oracle(cvr, oracle):
    sc = cvr.scenes.bucky()
    res, spp = 256, 16
    kl = cvr.NaiveVolPTsk(0)
    kl.setScene(sc)
    img = kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
    c = kl.counters()
    kl.close()
    cam = oracle.make_camera(res, res, res, res, fov_x=sc.fov_x)
    ref, oc = oracle.render_naive(_oscene(oracle, sc), cam, spp)
    ref = ref / spp
    assert c["paths"] == oc["paths"] == res * res * spp
    for k in ("bounces", "density_lookups", "albedo_lookups", "escaped"):
        assert abs(c[k] - oc[k]) / oc[k] <= 3e-3, (k, c[k], oc[k])
    rgb, rrgb = img[..., :3], ref[..., :3]
    assert float(np.sqrt(np.mean((rgb - rrgb) ** 2)) / rrgb.mean()) <= 0.01
    assert abs(float(rgb.mean()) - float(rrgb.mean())) / float(rrgb.mean()) <= 1e-3
    # alpha: 1/spp where at least one path escaped, identical pixel sets up to flipped paths
    assert np.mean((img[..., 3] > 0) == (ref[..., 3] > 0)) >= 0.999
    # same-seed pixels: the large majority identical to 1e-5
    assert np.mean(np.all(np.abs(rgb - rrgb) <= 1e-5, axis=2)) >= 0.85


def test_c2_hetvol_full_size_properties_and_sample(cvr, oracle):
    sc = cvr.scenes.hetvol()
    res, spp = 1024, 64
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    kl.setSeed(0)
    full = kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
    c = kl.counters()
    assert c["paths"] == res * res * spp
    # spp sharding over 8 ranks recomposes the image (same multiset of paths)
    acc = np.zeros_like(full)
    for r in range(8):
        kl.setSeed(0)
        acc += kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x, sample_first=8 * r, sample_count=8)
    good = ~np.isnan(full[..., :3]) & ~np.isnan(acc[..., :3])
    assert good.mean() > 0.9999
    assert np.max(np.abs(acc[..., :3][good] - full[..., :3][good])) <= 2e-5
    # oracle on a bounded sample: a 64x64 tile of the same 1024^2 image at 16 spp, same streams
    tile, off, s_spp = 64, (480, 480), 16
    cam = oracle.make_camera(tile, tile, res, res, off_x=off[0], off_y=off[1], fov_x=sc.fov_x)
    ref, _ = oracle.render_regen(_oscene(oracle, sc), cam, s_spp, seed=0, rng_mode=1)
    ref = ref[..., :3] / s_spp
    import torch

    d_tile = torch.zeros((tile, tile, 4), dtype=torch.float32, device="cuda:0")
    iv, rtv = cvr.abi.default_camera(res, res, sc.fov_x)
    kl.copyRasterToView(float(rtv[0]), float(rtv[1]))
    kl.setResolution(tile, tile)
    kl.copyPixelIndexRange(float(res), float(res))
    kl.copyInvViewMatrix(iv)
    kl.copyOffset(*off)
    kl.setNIterations(s_spp)
    kl.setSampleRange(0, 0)
    kl.setSeed(0)
    kl.setOutputPtr(d_tile.data_ptr())
    kl.launchRender()
    kl.sync()
    got = d_tile.cpu().numpy()[..., :3] / s_spp
    m = ~np.isnan(got)
    assert float(np.sqrt(np.mean((got[m] - ref[m]) ** 2)) / ref.mean()) <= 0.05
    kl.close()


def test_c3_manix_full_size_tiles(cvr, oracle):
    sc = cvr.scenes.manix()
    res, spp, tiles = 1024, 256, (10, 10)
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    kl.setSeed(0)
    host = np.full((res, res, 4), -3.0, np.float32)
    kl.renderImage((res, res), tiles, spp, fov_x=sc.fov_x, host_image=host, fuse_tiles=True)
    c = kl.counters()
    assert c["paths"] == 102 * 102 * 100 * spp  # Q6: 1020^2 pixels covered
    assert np.all(host[1020:] == -3.0) and np.all(host[:, 1020:] == -3.0)
    inner = host[:1020, :1020]
    assert np.nanmin(inner[..., :3]) >= 0.0 and 0.3 < np.nanmean(inner[..., :3]) < 1.0
    # 8-GPU tile sharding (tiles k = r mod 8) recomposes the same pixels
    img = np.full_like(host, -3.0)
    for r in range(8):
        kl.setSeed(0)
        kl.renderImage((res, res), tiles, spp, fov_x=sc.fov_x, host_image=img, tile_first=r, tile_stride=8,
                       fuse_tiles=True)
    a, b = inner[..., :3], img[:1020, :1020, :3]
    good = ~np.isnan(a) & ~np.isnan(b)
    assert np.max(np.abs(a[good] - b[good])) <= 2e-5
    assert np.all(img[1020:] == -3.0)
    # oracle on a bounded sample: tile 45 of the 10 x 10 table (origin 510, 408; 102 x 102) at 2 spp with
    # the stream base reset() leaves for that tile on one GPU (seed += n_paths per tile)
    rate = _oracle_sample(cvr, oracle, kl, _oscene(oracle, sc), sc.fov_x, (res, res), (102, 102), (510, 408), 2,
                          seed=(45 * 102 * 102 * 2) & 0xffffffff, variant=1, min_agree=0.97)
    print(f"C3 oracle sample: {rate:.4f} of 20808 paths within 1e-4")
    kl.close()


def test_more_than_2_to_32_paths_in_one_launch(cvr):
    """Q14: the reference's uint n_paths / int tid overflow beyond 2^32 paths per launch;
    here the path queue is 64-bit.  4096^2 pixels x 257 spp = 4.31e9 paths, camera looking
    away from the box so that every path escapes at once: the image must be exactly 1."""
    sc = cvr.scenes.bucky()
    res, spp = 4096, 257
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    img = kl.renderImage((res, res), (1, 1), spp, inv_view=[1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 100.0])
    c = kl.counters()
    assert c["paths"] == res * res * spp > 2 ** 32
    assert c["escaped"] == c["paths"] and c["density_lookups"] == 0
    assert np.array_equal(img[..., :3], np.ones((res, res, 3), np.float32))
    kl.close()


def test_hbm_resident_volume_runs_and_matches_small_grid_statistics(cvr):
    """fbm 384^3 with constant albedo 0.99 (C4-like, long multi-scatter paths): the cell8
    layout is 1.8 GB, far beyond L2, so lookups are HBM-bound; tracking=local must agree
    with tracking=global statistically."""
    sc = cvr.scenes.fbm(384)
    res, spp = 256, 16
    out = {}
    for tr in ("global", "local"):
        kl = cvr.RegenerationVolPTsk(0, tracking=tr)
        kl.setScene(sc)
        kl.setSeed(3)
        img = kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)[..., :3]
        out[tr] = (float(np.nanmean(img)), kl.counters())
        kl.close()
    (mg, cg), (ml, cl) = out["global"], out["local"]
    assert abs(mg - ml) / mg <= 0.01
    assert abs(cg["bounces"] - cl["bounces"]) / cg["bounces"] <= 0.02
    assert cl["density_lookups"] < cg["density_lookups"]


# ------------------------------------------------------------------ brick layout (sparse) and device-generated volumes
def _same_lookups(cvr, kl_a, kl_b, n=4096, seed=1):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-0.15, 1.15, size=(n, 3)).astype(np.float32)  # includes the wrap / clamp edges (Q2)
    p[:64] = rng.integers(0, 2, size=(64, 3)).astype(np.float32)  # exact corners
    da, _ = kl_a.debugLookup(p)
    db, _ = kl_b.debugLookup(p)
    return np.array_equal(da, db), float(np.abs(da - db).max())


def test_brick_layout_equals_dense_cells(cvr):
    """The sparse brick layout stores the SAME lookup cells as the dense cell8 layout: a
    device-generated sparse volume and the dense array of the same voxels give bit-identical
    lookups, event counters and images."""
    n, seed = 96, 4  # seed 4: 8 % of the 12^3 bricks of this small volume are active
    den, _, mx = cvr.abi.synth_volume("sparsefbm", n, n, n, seed, with_albedo=False)
    assert 0.005 < float((den > 0).mean()) < 0.5
    dense = cvr.Scene(den, None, (-0.5,) * 3, (0.5,) * 3, scale=100.0, max_density=mx, albedo_const=(0.99,) * 3)
    out = {}
    for name, sc in (("dense", dense), ("brick", cvr.scenes.sparse_fbm(n, seed))):
        kl = cvr.RegenerationVolPTsk(0)
        kl.setScene(sc)
        kl.setSeed(5)
        img = kl.renderImage((128, 128), (1, 1), 8, fov_x=0.7)
        out[name] = (kl, img, kl.counters(), kl.volumeInfo())
    assert out["brick"][3]["layout"] == "brick" and out["dense"][3]["layout"] == "cell8"
    assert 0 < out["brick"][3]["bricks"] < (n // 8 + 1) ** 3
    assert out["brick"][3]["layout_bytes"] < out["dense"][3]["layout_bytes"]
    same, err = _same_lookups(cvr, out["dense"][0], out["brick"][0])
    assert same, err
    for k in ("paths", "bounces", "density_lookups", "albedo_lookups", "escaped"):
        assert out["dense"][2][k] == out["brick"][2][k], k
    assert np.nanmax(np.abs(out["dense"][1] - out["brick"][1])) <= 5e-6
    # local-majorant tracking works on bricks too and skips the empty ones
    kl = cvr.RegenerationVolPTsk(0, tracking="local")
    kl.setScene(cvr.scenes.sparse_fbm(n, seed))
    kl.setSeed(5)
    img = kl.renderImage((128, 128), (1, 1), 8, fov_x=0.7)
    c = kl.counters()
    assert c["density_lookups"] < 0.5 * out["brick"][2]["density_lookups"]
    assert abs(float(np.nanmean(img[..., :3])) - float(np.nanmean(out["brick"][1][..., :3]))) <= 0.01
    kl.close()
    for v in out.values():
        v[0].close()


def test_device_generated_fbm_equals_host_generated(cvr):
    """cvr_set_scene_procedural("fbm") builds the dense cell8 layout on the device from the same
    noise functions the host generator uses (csrc/cvr_noise.h): identical lookups."""
    n = 64
    a = cvr.RegenerationVolPTsk(0)
    a.setScene(cvr.scenes.fbm(n))
    b = cvr.RegenerationVolPTsk(0)
    sc = cvr.scenes.fbm_device(n)
    b.setScene(sc)
    assert abs(sc.max_density - cvr.scenes.fbm(n).max_density) == 0.0
    same, err = _same_lookups(cvr, a, b)
    assert same, err
    a.close()
    b.close()


def test_vdb_leaves_to_bricks_equals_densified_grid(cvr, tmp_path):
    """VDB leaves re-laid into device bricks WITHOUT densifying (cvr_set_scene_sparse) give the
    lookups of the densified grid the reference builds (VDBAdapter.cpp:57-76) -- with a volume
    origin that is not a multiple of 8, so that bricks straddle leaves."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import vdb_writer as W

    rng = np.random.default_rng(4)
    shape = (45, 38, 52)
    d = rng.random(shape).astype(np.float32)
    z, y, x = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    blob = ((x - 20) ** 2 + (y - 18) ** 2 + (z - 25) ** 2 < 150) | ((x - 44) ** 2 + (y - 5) ** 2 + (z - 40) ** 2 < 60)
    d = np.where(blob, 0.2 + 0.8 * d, 0).astype(np.float32)
    d[0, 0, 0] = d[-1, -1, -1] = 0.5  # pin the bounding box
    path = REAL_VDB if os.path.exists(REAL_VDB) else None
    origin = (-13, 3, 21)
    if path is None:
        path = str(tmp_path / "s.vdb")
        a = np.stack([d, d, d], -1)
        W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0, origin)),
                           ("albedo", 3, (0, 0, 0), W.leaves_from_dense(a, d != 0, origin))])
    dense_sc = cvr.scenes.from_vdb(path)
    dense = cvr.Scene(dense_sc.density, None, dense_sc.box_min, dense_sc.box_max, scale=100.0,
                      max_density=dense_sc.max_density, albedo_const=(0.9, 0.8, 0.7))
    sparse = cvr.scenes.sparse_from_vdb(path, albedo_const=(0.9, 0.8, 0.7))
    ka, kb = cvr.RegenerationVolPTsk(0), cvr.RegenerationVolPTsk(0)
    ka.setScene(dense)
    kb.setScene(sparse)
    same, err = _same_lookups(cvr, ka, kb)
    assert same, err
    ka.setSeed(2), kb.setSeed(2)
    ia = ka.renderImage((96, 96), (1, 1), 8, fov_x=0.7)
    ib = kb.renderImage((96, 96), (1, 1), 8, fov_x=0.7)
    ca, cb = ka.counters(), kb.counters()
    for k in ("paths", "bounces", "density_lookups", "albedo_lookups", "escaped"):
        assert ca[k] == cb[k], k
    assert np.nanmax(np.abs(ia - ib)) <= 5e-6
    assert kb.volumeInfo()["layout"] == "brick" and kb.volumeInfo()["bricks"] > 0
    ka.close()
    kb.close()


def test_c4_fbm_1024_full_config(cvr, oracle):
    """C4: fBm 1024^3 (34.5 GB of lookup cells, generated on the device), albedo 0.99,
    2048 x 2048 x 128 spp = 5.4e8 paths in one launch; spp sharding recomposes at reduced spp."""
    sc = cvr.scenes.fbm_device(1024)
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    info = kl.volumeInfo()
    assert info["layout"] == "cell8" and info["layout_bytes"] == 1025 ** 3 * 32
    res, spp = 2048, 128
    kl.setSeed(0)
    img = kl.renderImage((res, res), (1, 1), spp, fov_x=sc.fov_x)
    c = kl.counters()
    assert c["paths"] == res * res * spp
    assert 0.3 < float(np.nanmean(img[..., :3])) < 0.99 and c["albedo_lookups"] / c["paths"] > 3.0  # long multi-scatter paths
    print(f"C4 fbm1024 2048^2 x {spp}: {c['paths'] / c['kernel_ms'] / 1e3:.1f} Msamples/s, "
          f"{c['density_lookups'] / c['paths']:.1f} lookups/path, {c['bounces'] / c['paths']:.2f} bounces/path")
    # sharding property at 512^2 x 8 spp
    kl.resetCounters()
    kl.setSeed(0)
    full = kl.renderImage((512, 512), (1, 1), 8, fov_x=sc.fov_x)
    acc = np.zeros_like(full)
    for r in range(4):
        kl.setSeed(0)
        acc += kl.renderImage((512, 512), (1, 1), 8, fov_x=sc.fov_x, sample_first=2 * r, sample_count=2)
    good = ~np.isnan(full[..., :3]) & ~np.isnan(acc[..., :3])
    assert np.max(np.abs(acc[..., :3][good] - full[..., :3][good])) <= 2e-5
    # oracle on a bounded sample: the SAME 1024^3 voxels generated on the host (4 GiB; the device
    # generator is pinned to it by test_device_generated_fbm_equals_host_generated), a 48 x 48 tile in
    # the middle of the 2048^2 image at 2 spp, constant albedo 0.99 as a 2x2x2 volume for the oracle
    den, _, mx = cvr.abi.synth_volume("fbm", 1024, 1024, 1024, 0, with_albedo=False)
    assert mx == sc.max_density
    osc = oracle.make_scene(den, _const_albedo((0.99,) * 3), sc.box_min, sc.box_max, sc.scale, mx)
    rate = _oracle_sample(cvr, oracle, kl, osc, sc.fov_x, (2048, 2048), (48, 48), (1000, 1000), 2, seed=0, variant=1,
                          min_agree=0.95)
    print(f"C4 oracle sample: {rate:.4f} of 4608 paths within 1e-4")
    del den, osc
    kl.close()


def test_c5_sparse_2048_config(cvr, oracle):
    """C5: sparse 2048^3 VDB-style volume in the brick layout (dense cells would be 275 GB),
    4096 x 4096; the full 1024 spp is 1.7e10 paths (the 64-bit path counter is covered by
    test_more_than_2_to_32_paths_in_one_launch), so the image properties are checked at 4 spp:
    tile and sample sharding over 8 ranks recompose the single-GPU image."""
    sc = cvr.scenes.sparse_fbm(2048)
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    info = kl.volumeInfo()
    assert info["layout"] == "brick" and 0.01 < info["bricks"] / 257 ** 3 < 0.25
    assert info["layout_bytes"] < 80e9
    res, spp = 4096, 4
    kl.setSeed(0)
    full = kl.renderImage((res, res), (8, 8), spp, fov_x=sc.fov_x, fuse_tiles=True)
    c = kl.counters()
    assert c["paths"] == res * res * spp
    assert c["albedo_lookups"] > 0.05 * c["paths"] and float(np.nanmean(full[..., :3])) < 0.999  # the medium is hit
    print(f"C5 sparse2048 4096^2 x {spp} (global majorant): {c['paths'] / c['kernel_ms'] / 1e3:.1f} Msamples/s, "
          f"{c['density_lookups'] / c['paths']:.1f} lookups/path, {c['bounces'] / c['paths']:.2f} bounces/path, "
          f"bricks {info['bricks']}, {info['layout_bytes'] / 1e9:.1f} GB")
    img = np.zeros_like(full)
    for r in range(8):  # tile sharding: disjoint pixels, seeds indexed by the global tile number
        kl.setSeed(0)
        kl.renderImage((res, res), (8, 8), spp, fov_x=sc.fov_x, host_image=img, tile_first=r, tile_stride=8, fuse_tiles=True)
    good = ~np.isnan(full[..., :3]) & ~np.isnan(img[..., :3])
    assert np.max(np.abs(img[..., :3][good] - full[..., :3][good])) <= 2e-5
    kl.close()
    # local-majorant tracking (skips the ~97 % empty bricks without a lookup): same image statistically
    kl = cvr.RegenerationVolPTsk(0, tracking="local")
    kl.setScene(sc)
    kl.setSeed(0)
    loc = kl.renderImage((res, res), (8, 8), spp, fov_x=sc.fov_x, fuse_tiles=True)
    cl = kl.counters()
    print(f"C5 sparse2048 4096^2 x {spp} (local majorants): {cl['paths'] / cl['kernel_ms'] / 1e3:.1f} Msamples/s, "
          f"{cl['density_lookups'] / cl['paths']:.1f} lookups/path")
    assert cl["density_lookups"] < 0.5 * c["density_lookups"]
    assert abs(float(np.nanmean(loc[..., :3])) - float(np.nanmean(full[..., :3]))) <= 0.005
    assert abs(cl["bounces"] - c["bounces"]) / c["bounces"] <= 0.02
    kl.close()
    # oracle on a brick-layout scene: the 2048^3 grid does not fit host memory as a dense array (32 GiB),
    # so the comparison runs on the same generator at 512^3 (512 MiB dense for the oracle, bricks on the
    # device), a 64 x 64 tile of the 4096^2 image at 2 spp
    n = 512
    ssc = cvr.scenes.sparse_fbm(n)
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(ssc)
    assert kl.volumeInfo()["layout"] == "brick"
    den, _, mx = cvr.abi.synth_volume("sparsefbm", n, n, n, 0, with_albedo=False)
    assert mx == ssc.max_density
    osc = oracle.make_scene(den, _const_albedo((0.99,) * 3), ssc.box_min, ssc.box_max, ssc.scale, mx)
    rate = _oracle_sample(cvr, oracle, kl, osc, ssc.fov_x, (res, res), (64, 64), (2000, 2100), 2, seed=0, variant=1,
                          min_agree=0.95)
    print(f"C5 oracle sample (sparse 512^3 bricks): {rate:.4f} of 8192 paths within 1e-4")
    kl.close()
