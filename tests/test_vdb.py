"""The OpenVDB-free .vdb reader (include/cvr_abi.h cvr_vdb_*, host/VdbReader.h) that replaces
the reference's VDBAdapter (implementation/vdb_adapter/VDBAdapter.{h,cpp}).

* pinned on the reference's own data file where the checkout is mounted (this container):
  the file's OpenVDB-written metadata (voxel count, bounding box, stream offsets), the
  converter's arithmetic (albedo = (density, 0, 0), values in [0, 1]) and committed digests;
* everywhere else on files written by tests/vdb_writer.py in every container / codec /
  node-mask-compression variant the reader supports;
* the error behaviour of VDBAdapter::loadVDBFile (VDBAdapter.cpp:15-43)."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import vdb_writer as W
from cudavolumerenderer_b200 import CvrError, scenes
from cudavolumerenderer_b200.vdb import VdbFile

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = "/root/reference/data/vdb/bonsai_small.vdb"


# ------------------------------------------------------------------ the real file
@pytest.mark.skipif(not os.path.exists(REAL), reason="reference checkout not mounted")
def test_real_bonsai_file_matches_its_own_metadata_and_golden():
    gold = json.load(open(os.path.join(HERE, "golden", "vdb_bonsai_small.json")))
    assert hashlib.sha256(open(REAL, "rb").read()).hexdigest() == gold["file_sha256"]
    with VdbFile(REAL) as f:
        grids = {g["name"]: g for g in f.grids()}
        assert set(grids) == {"density", "albedo"}
        assert grids["density"]["type"] == "Tree_float_5_4_3" and grids["albedo"]["type"] == "Tree_vec3s_5_4_3"
        dens = f.densify("density")
        alb = f.densify("albedo")
        for name, g in grids.items():
            # what OpenVDB itself recorded when it wrote the file
            assert g["active_voxels"] == int(f.meta(name, "file_voxel_count")) == 87684
            assert list(g["bbox_min"]) == [int(v) for v in f.meta(name, "file_bbox_min").split()] == [0, 0, 0]
            assert list(g["bbox_max"]) == [int(v) for v in f.meta(name, "file_bbox_max").split()] == [90, 196, 255]
            assert f.meta(name, "file_compression") == "blosc + active values" and g["compression"] == 6
            assert g["file_version"] == 224 and g["dim"] == (91, 197, 256)
        # the converter: density = smoothstep(...) in [0,1], albedo = (density, 0, 0)
        assert dens.shape == (256, 197, 91) and alb.shape == (256, 197, 91, 3)
        assert dens.min() == 0.0 and dens.max() == 1.0 and int((dens != 0).sum()) == 87684
        assert np.array_equal(alb[..., 0], dens) and not alb[..., 1:].any()
        # committed digests of the decoded values (tests/golden/make_vdb_golden.py)
        for name, dense in (("density", dens), ("albedo", alb)):
            gg = gold["grids"][name]
            assert hashlib.sha256(np.ascontiguousarray(dense).tobytes()).hexdigest() == gg["dense_sha256"]
            org, msk, val = f.leaves(name)
            assert len(org) == gg["info"]["leaf_count"] == 753
            assert hashlib.sha256(org.tobytes()).hexdigest() == gg["leaf_origin_sha256"]
            assert hashlib.sha256(msk.tobytes()).hexdigest() == gg["leaf_mask_sha256"]
            assert hashlib.sha256(np.ascontiguousarray(val).tobytes()).hexdigest() == gg["leaf_value_sha256"]
    sc = scenes.from_vdb(REAL)  # VDBSceneBuilder.h:40-80
    assert sc.density.shape == (256, 197, 91) and sc.albedo.shape == (256, 197, 91, 4)
    assert sc.scale == 100.0 and sc.max_density == 1.0 and tuple(sc.box_min) == (-0.5,) * 3
    assert np.all(sc.albedo[..., 3] == 1.0)


# ------------------------------------------------------------------ synthetic files
def _volume(seed=3, shape=(21, 30, 37)):
    """(nz, ny, nx) density with holes (inactive = exactly 0) and a matching vec3 albedo."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    d = np.sin(x * 0.31) * np.cos(y * 0.23) * np.sin(z * 0.41 + 1.0) + 0.15 * rng.standard_normal(shape)
    d = np.where(d > 0.1, d, 0.0).astype(np.float32)
    d[:, :, 0] = np.maximum(d[:, :, 0], 0.5)  # keep the bounding box at the array bounds
    d[-1, -1, -1] = 0.75
    d[0, 0, :] = np.maximum(d[0, 0, :], 0.25)
    d[:, 0, 0] = np.maximum(d[:, 0, 0], 0.25)
    a = np.stack([d, 0.5 * d, (d > 0) * 0.125], axis=-1).astype(np.float32)
    return d, a


@pytest.mark.parametrize("compression,blocksize,stored_every", [
    (W.COMPRESS_BLOSC | W.COMPRESS_ACTIVE_MASK, None, 0),  # what the reference's converter writes
    (W.COMPRESS_BLOSC | W.COMPRESS_ACTIVE_MASK, 512, 5),   # multi-block chunks, leftover blocks, stored chunks
    (W.COMPRESS_BLOSC, None, 0),                           # no active-mask compression: code 6 everywhere
    (W.COMPRESS_ZIP | W.COMPRESS_ACTIVE_MASK, None, 0),
    (W.COMPRESS_ZIP, None, 0),
    (W.COMPRESS_ACTIVE_MASK, None, 0),
    (0, None, 0),
])
def test_round_trip_of_written_files(tmp_path, compression, blocksize, stored_every):
    d, a = _volume()
    origin = (-11, 5, 130)  # negative coordinates, a bounding box that crosses 128-voxel node borders
    path = str(tmp_path / "t.vdb")
    W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0, origin)),
                       ("albedo", 3, (0, 0, 0), W.leaves_from_dense(a, d != 0, origin))],
                compression=compression, blosc_blocksize=blocksize, stored_every=stored_every, extra_grid=True)
    with VdbFile(path) as f:
        names = [g["name"] for g in f.grids()]
        assert names == ["ids", "density", "albedo"] and f.grids()[0]["channels"] == 0  # unknown grid skipped
        g = f.grid("density")
        assert g["bbox_min"] == origin and g["dim"] == (d.shape[2], d.shape[1], d.shape[0])
        assert g["active_voxels"] == int((d != 0).sum()) == int(f.meta("density", "file_voxel_count"))
        assert np.array_equal(f.densify("density"), d)
        assert np.array_equal(f.densify("albedo"), a)
        a4 = f.densify("albedo", out_channels=4, inactive=(0.5, 0.25, 0.125))
        assert np.array_equal(a4[..., :3][d != 0], a[d != 0]) and np.all(a4[..., 3] == 1.0)
        assert np.all(a4[..., :3][d == 0] == np.float32([0.5, 0.25, 0.125]))
        org, msk, val = f.leaves("density")
        assert len(org) == g["leaf_count"] and np.all(org % 8 == 0)
        # leaf value n <-> voxel origin + ((n>>6)&7, (n>>3)&7, n&7)
        n_on = 0
        for o, m, v in zip(org, msk, val):
            bits = np.unpackbits(m.view(np.uint8), bitorder="little").astype(bool)
            n_on += int(bits.sum())
            for n in np.flatnonzero(bits)[:7]:
                x, y, z = o[0] + (n >> 6) - origin[0], o[1] + ((n >> 3) & 7) - origin[1], o[2] + (n & 7) - origin[2]
                assert v[n] == d[z, y, x]
        assert n_on == g["active_voxels"]


def test_every_node_mask_compression_code(tmp_path):
    """Inactive voxels of a leaf are not stored under active-mask compression; a per-node code
    says how to rebuild them (0 +bg, 1 -bg, 2 one value, 3 mask -/+bg, 4 mask bg/value, 5 mask
    between two values, 6 all stored).  Background 0.5 so that -bg differs from +bg."""
    rng = np.random.default_rng(11)
    bg = np.float32(0.5)
    leaves, expect = [], []
    for k in range(7):
        mask = rng.random(512) < 0.4
        vals = rng.random(512).astype(np.float32) + 1.0
        ina = np.empty(512, np.float32)
        pick = rng.random(512) < 0.5
        if k == 0:
            ina[:] = bg
        elif k == 1:
            ina[:] = -bg
        elif k == 2:
            ina[:] = 7.0
        elif k == 3:
            ina[:] = np.where(pick, bg, -bg)
        elif k == 4:
            ina[:] = np.where(pick, bg, 3.0)
        elif k == 5:
            ina[:] = np.where(pick, 3.0, 4.0)
        else:
            ina[:] = rng.random(512) + 10.0
        vals[~mask] = ina[~mask]
        leaves.append(((8 * k, 0, 128 * (k % 2)), mask, vals.reshape(512, 1)))
        expect.append(vals)
    for comp in (W.COMPRESS_BLOSC | W.COMPRESS_ACTIVE_MASK, W.COMPRESS_ACTIVE_MASK, W.COMPRESS_ZIP | W.COMPRESS_ACTIVE_MASK):
        path = str(tmp_path / f"codes{comp}.vdb")
        W.write_vdb(path, [("density", 1, bg, leaves)], compression=comp)
        raw = open(path, "rb").read()
        with VdbFile(path) as f:
            assert f.grid("density")["background"][0] == 0.5
            org, msk, val = f.leaves("density")
            got = {tuple(o): v for o, v in zip(org, val)}
            for (o, m, _), e in zip(leaves, expect):
                assert np.array_equal(got[tuple(o)], e), (comp, o)
        assert len(raw) > 0


def test_half_float_grids(tmp_path):
    d, _ = _volume(seed=5, shape=(9, 12, 17))
    dh = d.astype(np.float16).astype(np.float32)
    dh[(d != 0) & (dh == 0)] = np.float32(np.float16(6e-5))  # keep the active set
    path = str(tmp_path / "h.vdb")
    W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(dh, d != 0))], half=True)
    with VdbFile(path) as f:
        assert np.array_equal(f.densify("density"), np.where(d != 0, dh, 0))


def test_lz4_long_runs_and_overlapping_matches(tmp_path):
    """Constant and periodic leaves compress to LZ4 sequences with 255-run length bytes and
    matches that overlap their own output."""
    vals = np.zeros((16, 16, 16), np.float32)
    vals[:8] = 0.625                                        # one constant leaf row
    vals[8:] = np.tile(np.float32([0.25, 0.5]), 8 * 16 * 8).reshape(8, 16, 16)
    path = str(tmp_path / "runs.vdb")
    W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(vals, vals != 0))])
    assert os.path.getsize(path) < 150000
    with VdbFile(path) as f:
        assert np.array_equal(f.densify("density"), vals)


# ------------------------------------------------------------------ errors (VDBAdapter.cpp:15-43)
def test_errors_are_reported(tmp_path):
    d, a = _volume(shape=(8, 8, 8))
    only_density = str(tmp_path / "d.vdb")
    W.write_vdb(only_density, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0))])
    with pytest.raises(CvrError, match="does not contain an albedo grid"):
        scenes.from_vdb(only_density)
    only_albedo = str(tmp_path / "a.vdb")
    W.write_vdb(only_albedo, [("albedo", 3, (0, 0, 0), W.leaves_from_dense(a, d != 0))])
    with pytest.raises(CvrError, match="does not contain a density grid"):
        scenes.from_vdb(only_albedo)
    with pytest.raises(CvrError, match="cannot open"):
        VdbFile(str(tmp_path / "missing.vdb"))
    bad = str(tmp_path / "bad.vdb")
    open(bad, "wb").write(b"not a vdb file at all, just text" * 4)
    with pytest.raises(CvrError, match="bad magic"):
        VdbFile(bad)
    raw = open(only_density, "rb").read()
    for cut in (len(raw) // 3, len(raw) - 7):
        open(bad, "wb").write(raw[:cut])
        with pytest.raises(CvrError):
            VdbFile(bad)
    old = bytearray(raw)
    old[8:12] = struct.pack("<I", 220)
    open(bad, "wb").write(bytes(old))
    with pytest.raises(CvrError, match="not supported"):
        VdbFile(bad)
    # a flipped byte inside a compressed stream must be an error or decode to different values, never a crash
    for pos in range(len(raw) - 400, len(raw) - 1, 37):
        mut = bytearray(raw)
        mut[pos] ^= 0x5A
        open(bad, "wb").write(bytes(mut))
        try:
            with VdbFile(bad) as f:
                f.densify("density")
        except CvrError:
            pass


# ------------------------------------------------------------------ C++ host layer (VDBSceneBuilder)
def _cli():
    import cudavolumerenderer_b200 as pkg

    cli = os.path.join(os.path.dirname(pkg.__file__), "cvr_render")
    if not os.path.exists(cli):
        pytest.skip("cvr_render not built")
    return cli


def test_cpp_vdb_scene_builder_matches_python_loader(tmp_path):
    """cvr_render <file>.vdb goes through the C++ VDBSceneBuilder (host/SceneBuilders.h over
    host/VdbReader.h); --dump-scene prints what it built, without touching a GPU."""
    import subprocess

    d, a = _volume(seed=9, shape=(13, 18, 22))
    path = str(tmp_path / "s.vdb")
    W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0, (3, -2, 40))),
                       ("albedo", 3, (0, 0, 0), W.leaves_from_dense(a, d != 0, (3, -2, 40)))])
    p = subprocess.run([_cli(), path, "--dump-scene"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr + p.stdout
    assert "Auto-detected scene type: Vdb" in p.stdout
    lines = {ln.split()[0]: ln.split()[1:] for ln in p.stdout.splitlines() if ln and ln.split()[0] in ("scene", "box", "scale", "sums")}
    sc = scenes.from_vdb(path)
    assert lines["scene"] == ["density", "22", "18", "13", "albedo", "22", "18", "13"]
    assert [float(v) for v in lines["box"]] == [-0.5, -0.5, -0.5, 0.5, 0.5, 0.5]
    assert float(lines["scale"][0]) == 100.0 and abs(float(lines["scale"][2]) - sc.max_density) < 1e-6
    assert abs(float(lines["sums"][0]) - float(sc.density.astype(np.float64).sum())) < 1e-3
    assert abs(float(lines["sums"][1]) - float(sc.albedo.astype(np.float64).sum())) < 1e-2
    # VDBAdapter's error for a file without albedo, surfaced by the CLI
    W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0))])
    p = subprocess.run([_cli(), path, "--dump-scene"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "does not contain an albedo grid" in p.stderr


@pytest.mark.gpu
def test_vdb_scene_renders_the_same_through_cli_and_python(tmp_path):
    import subprocess

    import cudavolumerenderer_b200 as cvr

    path = REAL
    if not os.path.exists(path):  # GPU box: no reference checkout -> a written file
        d, a = _volume(seed=2, shape=(40, 33, 29))
        path = str(tmp_path / "s.vdb")
        W.write_vdb(path, [("density", 1, 0.0, W.leaves_from_dense(d, d != 0)),
                           ("albedo", 3, (0, 0, 0), W.leaves_from_dense(a, d != 0))])
    raw = tmp_path / "img.bin"
    p = subprocess.run([_cli(), path, "-k", "naiveSK", "-r", "96", "-i", "4", "--interactive", "0", "--trials", "1",
                        "-o", str(tmp_path / "out"), "--dump-raw", str(raw)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    got = np.fromfile(raw, np.float32).reshape(96, 96, 4)
    sc = cvr.scenes.from_vdb(path)
    kl = cvr.NaiveVolPTsk(0)
    kl.setScene(sc)
    ref = kl.renderImage((96, 96), (1, 1), 4, fov_x=sc.fov_x)
    c = kl.counters()
    kl.close()
    assert np.allclose(got, ref, rtol=0, atol=2e-6)
    assert c["density_lookups"] > 0 and 0.0 < float(ref[..., :3].mean()) <= 1.0
