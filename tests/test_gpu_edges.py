"""Edge cases of the path on the GPU, each checked against the CPU oracle (same seeds, naiveSK
semantics = Rng(path id)) or an exact property: degenerate image and volume sizes, tiles
that do not divide the resolution, fewer paths than one warp, an empty medium, an opaque one,
non-cubic grids and boxes, bounce caps, roulette off."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cvr():
    import cudavolumerenderer_b200 as pkg

    return pkg


def _scene(cvr, den, alb=None, box=((-0.5,) * 3, (0.5,) * 3), scale=20.0, mx=None):
    den = np.ascontiguousarray(den, np.float32)
    if alb is None:
        alb = np.empty(den.shape + (4,), np.float32)
        alb[..., 0], alb[..., 1], alb[..., 2], alb[..., 3] = 0.9, 0.7, 0.5, 1.0
    return cvr.Scene(den, alb, box[0], box[1], scale=scale, max_density=float(den.max()) if mx is None else mx)


def _oracle_image(oracle, sc, res, spp, tile=None, off=(0, 0)):
    osc = oracle.make_scene(sc.density, sc.albedo, sc.box_min, sc.box_max, sc.scale, sc.max_density)
    tw, th = tile or res
    cam = oracle.make_camera(tw, th, res[0], res[1], off_x=off[0], off_y=off[1], fov_x=sc.fov_x)
    img, ctr = oracle.render_naive(osc, cam, spp)
    return img / spp, ctr


def _close(got, ref, frac=0.97):
    """same seeds: all but the ulp-flipped paths agree; at low spp that is almost every pixel"""
    ok = np.all(np.abs(got[..., :3] - ref[..., :3]) <= 1e-4, axis=-1)
    return ok.mean() >= frac


@pytest.mark.parametrize("res,tiles,spp", [((1, 1), (1, 1), 1), ((3, 2), (1, 1), 5), ((7, 5), (3, 2), 3),
                                           ((33, 17), (4, 3), 2), ((5, 64), (5, 1), 1)])
@pytest.mark.parametrize("exact", [0, 1])
def test_small_and_ragged_images_match_the_oracle(cvr, oracle, res, tiles, spp, exact):
    """Fewer paths than one warp, tiles that do not divide the resolution (the remainder is
    never rendered, Q6), one-pixel tiles: naiveSK seeds every tile with the bare path id."""
    sc = cvr.scenes.bucky()
    kl = cvr.NaiveVolPTsk(0, exact=exact)
    kl.setScene(sc)
    host = np.full((res[1], res[0], 4), -7.0, np.float32)
    kl.renderImage(res, tiles, spp, fov_x=sc.fov_x, host_image=host)
    c = kl.counters()
    kl.close()
    dim, org = cvr.abi.tile_table(res[0], res[1], tiles[0], tiles[1])
    tw, th = int(dim[0]), int(dim[1])
    assert c["paths"] == tw * th * tiles[0] * tiles[1] * spp
    covered = np.zeros((res[1], res[0]), bool)
    for ox, oy in org:
        ref, _ = _oracle_image(oracle, sc, res, spp, tile=(tw, th), off=(int(ox), int(oy)))
        got = host[oy:oy + th, ox:ox + tw]
        assert _close(got, ref, 0.9 if tw * th < 64 else 0.97), (res, tiles, ox, oy)
        covered[oy:oy + th, ox:ox + tw] = True
    assert np.all(host[~covered] == -7.0)  # untouched remainder


def test_empty_medium_is_two_boundary_events_per_path(cvr, oracle):
    """density == 0 everywhere: every path enters and leaves through the rough-dielectric
    boundary without a single collision; same-seed agreement with the oracle."""
    den = np.zeros((6, 5, 4), np.float32)
    sc = _scene(cvr, den, mx=1.0)
    kl = cvr.NaiveVolPTsk(0)
    kl.setScene(sc)
    img = kl.renderImage((40, 40), (1, 1), 4, fov_x=sc.fov_x)
    c = kl.counters()
    kl.close()
    ref, oc = _oracle_image(oracle, sc, (40, 40), 4)
    assert c["albedo_lookups"] == 0 == oc["albedo_lookups"]
    assert c["paths"] == oc["paths"] and abs(c["bounces"] - oc["bounces"]) <= 0.002 * oc["bounces"]
    assert _close(img, ref)
    assert float(np.nanmean(img[..., :3])) > 0.9


def test_opaque_absorbing_medium(cvr, oracle):
    """albedo 0: the first collision kills the path (roulette survival 0, Q9)."""
    den = np.ones((8, 8, 8), np.float32)
    alb = np.zeros((8, 8, 8, 4), np.float32)
    alb[..., 3] = 1.0
    sc = _scene(cvr, den, alb, scale=400.0)
    kl = cvr.RegenerationVolPTsk(0, rng="xorwow-path")
    kl.setScene(sc)
    kl.setSeed(0)
    img = kl.renderImage((32, 32), (1, 1), 8, fov_x=sc.fov_x)
    c = kl.counters()
    kl.close()
    hit = img[8:24, 8:24, :3]
    assert float(np.nanmax(hit)) < 0.2          # almost nothing gets through the box
    assert c["albedo_lookups"] <= c["paths"]      # at most one collision per path
    assert np.all(img[0, :, :3] == 1.0)           # rays that miss the box: exactly the environment


@pytest.mark.parametrize("shape,box", [((2, 2, 2), ((-0.5,) * 3, (0.5,) * 3)),
                                       ((3, 50, 7), ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))),
                                       ((9, 4, 31), ((-0.64, -0.3, -0.25), (0.64, 0.3, 0.25)))])
def test_minimal_and_non_cubic_grids_match_the_oracle(cvr, oracle, shape, box):
    """2^3 is the smallest grid with a trilinear cell; non-cubic grids and non-unit boxes
    exercise the Q1 coordinate quirk (density is sampled at p + 0.5 whatever the box)."""
    rng = np.random.default_rng(sum(shape))
    den = rng.random(shape).astype(np.float32)
    sc = _scene(cvr, den, box=box, scale=30.0)
    for exact in (1, 0):
        kl = cvr.NaiveVolPTsk(0, exact=exact)
        kl.setScene(sc)
        img = kl.renderImage((48, 40), (1, 1), 4, fov_x=sc.fov_x)
        c = kl.counters()
        kl.close()
        ref, oc = _oracle_image(oracle, sc, (48, 40), 4)
        assert _close(img, ref), (shape, exact)
        assert abs(c["density_lookups"] - oc["density_lookups"]) <= 0.01 * oc["density_lookups"]


def test_bounce_cap_and_roulette_off(cvr):
    """max_bounces ends every path after that many loop iterations; with the roulette off no
    path dies before the cap or an escape, and the parked boundary uniform of the fused
    tracking loop is handed back when nobody consumes it (Xorwow::undo)."""
    sc = cvr.scenes.bucky()
    res, spp = (64, 64), 4
    for sched in ("warp", "queued", "lane"):
        kl = cvr.createLauncher("regenerationSK", 0, sched=sched, max_bounces=3)
        kl.setScene(sc)
        kl.setSeed(1)
        kl.renderImage(res, (1, 1), spp, fov_x=sc.fov_x)
        c = kl.counters()
        assert c["bounces"] <= 3 * c["paths"], sched
        kl.close()
    out = {}
    for sched, exact in (("lane", 1), ("warp", 1), ("warp", 0), ("queued", 0)):
        kl = cvr.createLauncher("regenerationSK", 0, sched=sched, exact=exact, russian_roulette=0, max_bounces=64)
        kl.setScene(sc)
        kl.setSeed(1)
        img = kl.renderImage(res, (1, 1), spp, fov_x=sc.fov_x)
        out[(sched, exact)] = (img, kl.counters())
        kl.close()
    a, b = out[("lane", 1)], out[("warp", 1)]
    for k in ("paths", "bounces", "density_lookups", "albedo_lookups", "escaped"):
        assert a[1][k] == b[1][k], k
        assert out[("warp", 0)][1][k] == out[("queued", 0)][1][k], k
    assert np.nanmax(np.abs(a[0] - b[0])) <= 5e-6
    # fused vs reference-order arithmetic: the same paths up to ulp-flipped ones
    assert abs(out[("warp", 0)][1]["bounces"] - a[1]["bounces"]) <= 0.003 * a[1]["bounces"]


def test_zero_iterations_and_repeated_scene_changes(cvr):
    sc = cvr.scenes.bucky()
    kl = cvr.RegenerationVolPTsk(0)
    kl.setScene(sc)
    with pytest.raises(cvr.CvrError):
        kl.renderImage((16, 16), (1, 1), 0, fov_x=sc.fov_x)
    # scene changes between dense, sparse and procedural layouts on one handle
    imgs = []
    for s in (sc, cvr.scenes.sparse_fbm(64, 4), cvr.scenes.hetvol(), cvr.scenes.fbm_device(32), sc):
        kl.setScene(s)
        kl.setSeed(3)
        imgs.append(kl.renderImage((32, 32), (1, 1), 2, fov_x=0.7))
    assert np.nanmax(np.abs(imgs[0] - imgs[-1])) <= 5e-6
    kl.close()
