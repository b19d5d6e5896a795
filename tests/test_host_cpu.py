"""CPU tests of the host layer: the C++ scene loaders behind cvr_render (Raw, Mitsuba
XML + VOL v3, Vdb error path, flag parsing) and the multi-GPU sharding rules, including
a world_size-2 gloo run of the reduction logic."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cudavolumerenderer_b200", "cvr_render")


@pytest.fixture(scope="module")
def cli():
    import __graft_entry__ as g

    if not os.path.exists(CLI):
        g.build()
    return CLI


def run(cli, *args):
    p = subprocess.run([cli, *args], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout, p.stderr


def parse_dump(out):
    d = {}
    for line in out.splitlines():
        t = line.split()
        if t and t[0] in ("scene", "box", "scale", "sums"):
            d[t[0]] = t[1:]
    return d


def test_raw_loader_matches_reference_arithmetic(cli, tmp_path):
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 200, 32 ** 3, dtype=np.uint8)
    f = tmp_path / "vol.raw"
    raw.tofile(f)
    rc, out, err = run(cli, str(f), "--dump-scene")
    assert rc == 0, err
    d = parse_dump(out)
    assert d["scene"] == ["density", "32", "32", "32", "albedo", "32", "32", "32"]
    assert [float(x) for x in d["box"]] == [-0.5] * 3 + [0.5] * 3
    assert float(d["scale"][0]) == 40.0 and float(d["scale"][2]) == 1.0  # scale, max_density
    den = raw.astype(np.float32) / np.float32(raw.max())  # RawSceneBuilder.h:54-64
    assert abs(float(d["sums"][0]) - float(den.astype(np.float64).sum())) < 1e-2
    assert "Auto-detected scene type: Raw" in out
    # wrong size is an error, not a crash
    (tmp_path / "short.raw").write_bytes(b"\0" * 100)
    rc, out, err = run(cli, str(tmp_path / "short.raw"), "--dump-scene")
    assert rc == 1 and "Error" in err


def test_python_scenes_go_through_the_cpp_loaders(cli, tmp_path):
    """cvr_scene_file_load (include/cvr_abi.h): the ctypes layer loads scenes through the C++ SceneBuilders of
    cvr_render -- one implementation of the loaders.  Raw file: the exact voxels RawSceneBuilder.h:54-64 produces;
    stand-ins: the spec grammar, shapes and medium parameters; errors carry the loader's message."""
    from cudavolumerenderer_b200 import abi, scenes

    rng = np.random.default_rng(9)
    raw = rng.integers(0, 200, 32 ** 3, dtype=np.uint8)
    f = tmp_path / "vol.raw"
    raw.tofile(f)
    sc = scenes.load(str(f))
    den = raw.astype(np.float32) / np.float32(raw.max())
    assert sc.density.shape == (32, 32, 32) and np.array_equal(sc.density.ravel(), den)
    assert sc.albedo.shape == (32, 32, 32, 4) and np.all(sc.albedo[..., 3] == 1.0)
    assert (sc.scale, sc.max_density, sc.box_min, sc.box_max) == (40.0, 1.0, (-0.5,) * 3, (0.5,) * 3)
    assert abs(sc.fov_x - 0.7) < 1e-7 and abs(sc.ggx_eta - float(np.float32(1.05) / np.float32(1.01))) < 1e-7
    info = abi.load_scene_file(str(f), "Raw")
    assert info["type"] == "Raw" and info["resolution"] == (400, 400)
    # the stand-ins: same voxels as the generator called directly, parameters from the ONE table in SceneBuilders.h
    het = scenes.hetvol(dims=(24, 20, 10), seed=3)
    d0, a0, mx = abi.synth_volume("hetvol", 24, 20, 10, 3)
    assert np.array_equal(het.density, d0) and np.array_equal(het.albedo, a0) and het.max_density == mx
    f32 = lambda t: tuple(float(np.float32(v)) for v in t)  # the medium box travels as fp32 (cvr_scene_desc)
    assert (het.scale, het.box_min, het.box_max) == (800.0, f32((-0.64, -0.64, -0.25)), f32((0.64, 0.64, 0.25)))
    assert abs(het.fov_x - 0.33) < 1e-7
    fb = scenes.fbm(16)
    assert fb.albedo is None and fb.albedo_const == (0.99, 0.99, 0.99) and fb.density.shape == (16, 16, 16) and fb.name == "fbm16"
    assert abs(abi.load_scene_file("synth:fbm:16")["albedo_const"][0] - 0.99) < 1e-7
    assert scenes.manix(dims=(12, 10, 8)).density.shape == (8, 10, 12) and scenes.bucky().density.shape == (32, 32, 32)
    for bad, msg in ((str(tmp_path / "missing.raw"), "Error opening file"), ("synth:nosuch", "unknown synthetic scene"),
                     ("synth:fbm:abc", "cannot parse"), (str(tmp_path / "missing.vdb"), "OpenVDB error")):
        with pytest.raises(abi.CvrError, match=msg):
            scenes.load(bad)
    with pytest.raises(abi.CvrError, match="scene type not correct"):
        scenes.load(str(f), "Obj")


def write_vol(path, data, box):
    nz, ny, nx = data.shape[:3]
    ch = 1 if data.ndim == 3 else data.shape[3]
    with open(path, "wb") as f:
        f.write(b"VOL" + bytes([3]) + struct.pack("<i3ii", 1, nx, ny, nz, ch) + struct.pack("<6f", *box))
        f.write(np.ascontiguousarray(data, np.float32).tobytes())


def test_xml_vol_loader_reproduces_reference_quirks(cli, tmp_path):
    rng = np.random.default_rng(3)
    den = (rng.random((5, 6, 7)) * 1.5).astype(np.float32)  # values above 1: max_density clamps to 1
    alb = rng.random((5, 6, 7, 3)).astype(np.float32)
    write_vol(tmp_path / "smoke.vol", den, (-0.5, -0.5, -0.2, 0.5, 0.5, 0.2))
    write_vol(tmp_path / "albedo.vol", alb, (-0.64, -0.64, -0.25, 0.64, 0.64, 0.25))
    (tmp_path / "scene.xml").write_text("""<?xml version="1.0"?>
<scene version="0.5.0">
 <medium type="heterogeneous" id="smoke">
  <string name="method" value="woodcock"/>
  <volume name="density" type="gridvolume"><string name="filename" value="smoke.vol"/></volume>
  <volume name="albedo" type="gridvolume"><string name="filename" value="albedo.vol"/></volume>
  <float name="scale" value="800"/>
 </medium>
 <sensor type="perspective"><float name="fov" value="0.33"/>
  <film type="hdrfilm"><integer name="height" value="400"/><integer name="width" value="400"/></film>
 </sensor>
</scene>""")
    rc, out, err = run(cli, str(tmp_path / "scene.xml"), "--dump-scene", "-r", "512")
    assert rc == 0, err
    d = parse_dump(out)
    assert d["scene"] == ["density", "7", "6", "5", "albedo", "7", "6", "5"]
    # Q3: the medium box is the ALBEDO file's box
    assert np.allclose([float(x) for x in d["box"]], [-0.64, -0.64, -0.25, 0.64, 0.64, 0.25])
    assert float(d["scale"][0]) == 800.0
    assert float(d["scale"][2]) == 1.0  # max(min(1, v)), XmlSceneBuilder.h:187
    assert abs(float(d["scale"][4]) - 0.33) < 1e-6
    # albedo alpha = 1 per voxel
    assert abs(float(d["sums"][1]) - (float(alb.astype(np.float64).sum()) + 5 * 6 * 7)) < 1e-2
    assert "Auto-detected scene type: MitsubaXml" in out
    # the same file through the C ABI (cvr_scene_file_load) = what the ctypes layer hands to cvr_set_scene: same loader, same quirks
    from cudavolumerenderer_b200 import abi, scenes

    sc = scenes.load(str(tmp_path / "scene.xml"))
    assert sc.density.shape == (5, 6, 7) and sc.albedo.shape == (5, 6, 7, 4)
    assert np.array_equal(sc.density, den) and np.array_equal(sc.albedo[..., :3], alb) and np.all(sc.albedo[..., 3] == 1.0)
    assert np.allclose(sc.box_min + sc.box_max, [-0.64, -0.64, -0.25, 0.64, 0.64, 0.25]) and sc.scale == 800.0 and sc.max_density == 1.0
    info = abi.load_scene_file(str(tmp_path / "scene.xml"), "MitsubaXml")
    assert info["type"] == "MitsubaXml" and abs(info["fov_x"] - 0.33) < 1e-6
    bad = tmp_path / "bad.vol"
    bad.write_bytes(b"XXX" + b"\0" * 64)
    (tmp_path / "bad.xml").write_text((tmp_path / "scene.xml").read_text().replace("smoke.vol", "bad.vol"))
    rc, out, err = run(cli, str(tmp_path / "bad.xml"), "--dump-scene")
    assert rc == 1 and "incorrect header identifier" in err
    with pytest.raises(abi.CvrError, match="incorrect header identifier"):
        scenes.load(str(tmp_path / "bad.xml"))


def test_cli_flags_and_error_paths(cli, tmp_path):
    rc, out, err = run(cli, "--help")
    assert rc == 0 and "--number-of-tiles" in out and "--kernel" in out
    rc, out, err = run(cli)
    assert rc == 1 and "no scene file provided" in err
    rc, out, err = run(cli, "x.vdb", "--dump-scene")
    assert rc == 1 and "OpenVDB" in err
    rc, out, err = run(cli, "synth:bucky", "--dump-scene", "-k", "naiveSK", "-i", "16", "-r", "256", "--number-of-tiles", "2", "3")
    assert rc == 0 and "kernel set to naiveSK" in out and "iterations set to 16" in out
    rc, out, err = run(cli, "synth:bucky", "-a", "bidir", "--dump-scene")
    assert rc == 1 and "algorithm" in err
    rc, out, err = run(cli, "synth:bucky", "--bogus")
    assert rc == 1 and "unrecognised option" in err


def test_sharding_rules():
    from cudavolumerenderer_b200.distributed import spp_shard, tile_shard

    for total in (64, 65, 7, 1024):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                first, count = spp_shard(total, r, world)
                cover += list(range(first, first + count))
            assert cover == list(range(total))
    for n in (1, 9, 100):
        for world in (1, 2, 8):
            tiles = sorted(t for r in range(world) for t in tile_shard(n, r, world))
            assert tiles == list(range(n))
    with pytest.raises(ValueError):
        spp_shard(8, 2, 2)


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CVR_ROOT"])
from cudavolumerenderer_b200.distributed import render_sharded, spp_shard, tile_shard

class FakeLauncher:
    # stands in for the CUDA launcher: 'renders' a deterministic function of (pixel, sample)
    def renderImage(self, res, n_tiles, iterations, fov_x=0.7, inv_view=None, tile_first=0, tile_stride=1,
                    sample_first=0, sample_count=0, fuse_tiles=False, host_image=None, d_image=None):
        W, H = res
        img = IMG
        count = sample_count or iterations
        ys, xs = np.mgrid[0:H, 0:W]
        tw, th = W // n_tiles[0], H // n_tiles[1]
        tile_id = (ys // th) * n_tiles[0] + (xs // tw)
        covered = (xs < tw * n_tiles[0]) & (ys < th * n_tiles[1])
        mine = covered & (tile_id % tile_stride == tile_first % tile_stride) if tile_stride > 1 else covered
        acc = np.zeros((H, W), np.float64)
        for s in range(sample_first, sample_first + count):
            acc += ((xs * 31 + ys * 17 + s * 7) % 13) / 13.0
        img[..., 0] += torch.from_numpy(np.where(mine, acc / iterations, 0.0).astype(np.float32))
        # alpha as the kernels leave it: an escaping path STORES w = 1, the resolve divides by the iteration count
        img[..., 3] += torch.from_numpy(np.where(mine & (count > 0), np.float32(1) / np.float32(iterations), 0).astype(np.float32))

    def renderImageSharded(self, res, n_tiles, iterations, shard, fov_x=0.7, inv_view=None, fuse_tiles=True,
                           host_image=None, d_image=None):
        # the two parts of a cvr_shard: whole tiles with every sample, tail tiles with this rank's samples
        W, H = res
        ys, xs = np.mgrid[0:H, 0:W]
        tw, th = W // n_tiles[0], H // n_tiles[1]
        tile_id = (ys // th) * n_tiles[0] + (xs // tw)
        covered = (xs < tw * n_tiles[0]) & (ys < th * n_tiles[1])
        whole = covered & (tile_id >= shard.tile_first) & (tile_id < shard.tile_limit) & \
            ((tile_id - shard.tile_first) % max(shard.tile_stride, 1) == 0)
        tail = covered & (tile_id >= shard.tail_first) & (tile_id < shard.tail_limit)
        acc = np.zeros((H, W), np.float64)
        for s in range(iterations):
            v = ((xs * 31 + ys * 17 + s * 7) % 13) / 13.0
            in_tail = shard.sample_first <= s < shard.sample_first + shard.sample_count
            acc += np.where(whole | (tail & in_tail), v, 0.0)
        IMG[..., 0] += torch.from_numpy((acc / iterations).astype(np.float32))
        seen = whole | (tail & (shard.sample_count > 0))
        IMG[..., 3] += torch.from_numpy(np.where(seen, np.float32(1) / np.float32(iterations), 0).astype(np.float32))

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["CVR_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
out = {}
for mode in ("spp", "tiles", "balanced"):
    IMG = torch.zeros((12, 16, 4), dtype=torch.float32)
    render_sharded(FakeLauncher(), (16, 12), (2, 2), 10, mode, IMG)
    out[mode] = IMG.numpy().copy()
# 3 x 3 tiles on 2 ranks: 8 whole tiles + 1 tail tile split by sample index, reduced to rank 0 only
IMG = torch.zeros((12, 18, 4), dtype=torch.float32)
render_sharded(FakeLauncher(), (18, 12), (3, 3), 7, "balanced", IMG, reduce_to=0)
out["balanced33"] = IMG.numpy().copy()
if dist.get_rank() == 0:
    np.savez(os.environ["CVR_OUT"], **out)
dist.barrier()
dist.destroy_process_group()
"""


def test_world_size_2_gloo_reduction(tmp_path):
    """render_sharded on two CPU ranks over gloo: both sharding modes reproduce the
    single-process image after the one all-reduce."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = tmp_path / "worker.py"
    worker.write_text(GLOO_WORKER)
    outs = {}
    for world in (1, 2):
        procs = []
        outp = tmp_path / f"out{world}.npz"
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), CVR_PORT=str(port + world), CVR_ROOT=ROOT,
                       CVR_OUT=str(outp))
            procs.append(subprocess.Popen([sys.executable, str(worker)], env=env, stderr=subprocess.PIPE, text=True))
        for p in procs:
            _, err = p.communicate(timeout=180)
            assert p.returncode == 0, err[-2000:]
        outs[world] = np.load(outp)
    for mode in ("spp", "tiles", "balanced", "balanced33"):
        assert np.allclose(outs[1][mode], outs[2][mode], atol=1e-6), mode
        assert outs[1][mode][..., 0].sum() > 0
        # alpha is not a sum over the sample-sharded ranks: 1 / iterations wherever a path was traced, exactly
        assert np.array_equal(outs[1][mode][..., 3], outs[2][mode][..., 3]), mode
        assert set(np.unique(outs[2][mode][..., 3])) <= {np.float32(0), np.float32(1) / np.float32(7 if mode == "balanced33" else 10)}
    assert np.allclose(outs[1]["spp"], outs[1]["balanced"], atol=1e-6)


def test_shard_plans_cover_every_tile_sample_pair_once():
    """cvr_shard_plan (C ABI): over all ranks every (tile, sample) pair is rendered exactly once in
    every mode, and the balanced mode gives every rank the same amount of work up to one sample of
    the tail tiles (100 tiles x 250 samples on 8 ranks: 12 whole tiles + a share of 4 tiles each; 256 samples:
    32 samples of every tile each)."""
    from cudavolumerenderer_b200 import abi

    for n_tiles, iters in ((100, 256), (1, 64), (64, 16), (9, 7), (5, 3), (7, 1)):
        for world in (1, 2, 3, 4, 8):
            for mode in ("tiles", "spp", "balanced"):
                cover = np.zeros((n_tiles, iters), np.int32)
                work = []
                for r in range(world):
                    sh = abi.shard_plan(n_tiles, iters, r, world, mode)
                    w = 0
                    for k in range(sh.tile_first, min(sh.tile_limit, n_tiles), max(sh.tile_stride, 1)):
                        cover[k, :] += 1
                        w += iters
                    for k in range(sh.tail_first, sh.tail_limit):
                        cover[k, sh.sample_first:sh.sample_first + sh.sample_count] += 1
                        w += sh.sample_count
                    work.append(w)
                assert np.all(cover == 1), (n_tiles, iters, world, mode)
                if mode == "balanced":
                    assert max(work) - min(work) <= n_tiles % world, (n_tiles, iters, world, work)
    # samples divide evenly -> the sample split (same paths on every rank, one launch) ...
    sh = abi.shard_plan(100, 256, 3, 8, "balanced")
    assert (sh.tile_first, sh.tile_limit, sh.tail_first, sh.tail_limit, sh.sample_first, sh.sample_count) == (0, 0, 0, 100, 96, 32)
    # ... otherwise whole rounds of tiles by rank and the left-over tiles by sample index
    sh = abi.shard_plan(100, 250, 3, 8, "balanced")
    assert (sh.tile_first, sh.tile_stride, sh.tile_limit, sh.tail_first, sh.tail_limit, sh.sample_first, sh.sample_count) == \
        (3, 8, 96, 96, 100, 95, 31)
    with pytest.raises(ValueError):
        abi.shard_plan(10, 4, 2, 2, "balanced")
    with pytest.raises(ValueError):
        abi.shard_plan(10, 4, 0, 2, "rows")


def test_fetch_skip_table_quantisation_is_conservative():
    """The byte the fetch-skip table stores per brick (k_build_skip_table, cvr_kernels.cuh):
    s = clamp(ceil(r * 256 * (1 + 1e-5)), 1, 256) - 1 with r = majorant * sig_ratio, and the kernel
    skips a cell load when (word >> 24) > s.  Restated here in fp32 and checked exhaustively: whenever
    the test fires, the accept draw w = word * 2^-32 + 2^-33 (curand_uniform) is strictly above
    r * (1 + 4 ulp) -- the most a fused trilinear blend of corners <= majorant can reach -- so the
    skipped step is a null collision for certain; s = 255 never skips; r > 1 or NaN never skips."""
    f32 = np.float32

    def table_byte(r):
        r = f32(r)
        q = np.ceil(f32(f32(r * f32(256.0)) * f32(1.00001)))
        q = f32(1.0) if not (q >= 1.0) else (f32(256.0) if q > 256.0 else q)
        if not (r <= f32(1.0)):
            q = f32(256.0)
        return int(q) - 1

    rng = np.random.default_rng(7)
    rs = np.concatenate([np.linspace(0, 1, 4097), (np.arange(257) / 256.0), (np.arange(1, 257) / 256.0) * (1 - 2e-7),
                         rng.random(20000)]).astype(np.float32)
    for r in rs:
        s = table_byte(r)
        assert 0 <= s <= 255
        if s == 255:
            continue  # top byte is at most 255: the test `top > 255` never fires
        top = s + 1  # the smallest top byte that skips; w is smallest for word = top << 24
        w_min = f32(f32(np.uint32(top << 24)) * f32(2.3283064e-10) + f32(1.1641532e-10))
        bound = f32(r) * f32(1 + 4 * 1.1920929e-07)
        assert w_min > bound, (float(r), s, float(w_min), float(bound))
    assert table_byte(0.0) == 0 and table_byte(1.0) == 255 and table_byte(1.5) == 255 and table_byte(float("nan")) == 255
    assert table_byte(0.5) == 128  # ceil(128.00128) - 1: one step of margin above an exactly representable r
