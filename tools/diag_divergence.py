"""Diagnostic: three-way per-path comparison ours / reference GPU kernel / CPU oracle on bucky at 1 spp
(each pixel = one path), and details of paths that differ."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cudavolumerenderer_b200 as cvr
from oracle import bindings as oracle
import test_gpu_parity as T
import parity_util as PU

res = 128
for scene in ("bucky", "hetvol"):
    sc = cvr.scenes.make(scene)
    R = T._ref_gpu()
    T._ref_gpu_set_scene(R, sc)
    iv, rtv = cvr.abi.default_camera(res, res, sc.fov_x)
    ref1, _ = T._ref_gpu_render(R, 0, (res, res), (res, res), (0, 0), 1, 0, iv, rtv)
    R.refgpu_release()
    osc = oracle.make_scene(sc.density, sc.albedo, sc.box_min, sc.box_max, sc.scale, sc.max_density)
    cam = oracle.make_camera(res, res, res, res, fov_x=sc.fov_x)
    n = res * res
    ora, _ = oracle.trace_paths_naive(osc, cam, 0, n)
    for exact in (0, 1):
        kl = cvr.NaiveVolPTsk(0, exact=exact)
        kl.setScene(sc)
        got, log = T._trace_tile(cvr, kl, sc, (res, res), (res, res), (0, 0), 1, log_cap=256)
        kl.close()
        r = ref1.reshape(n, 4)
        a_or = np.all(np.abs(got[:, :3] - ora[:, :3]) <= 1e-4, axis=1)
        a_rg = np.all(np.abs(got[:, :3] - r[:, :3]) <= 1e-4, axis=1)
        o_rg = np.all(np.abs(ora[:, :3] - r[:, :3]) <= 1e-4, axis=1)
        print(scene, "exact", exact, "ours~oracle", a_or.mean(), "ours~refgpu", a_rg.mean(), "oracle~refgpu", o_rg.mean(),
              "ours==refgpu bitwise", float(np.all(got[:, :3] == r[:, :3], axis=1).mean()),
              "ours==oracle bitwise", float(np.all(got[:, :3] == ora[:, :3], axis=1).mean()))
        bad = np.nonzero(~a_or)[0]
        shown = 0
        for p in bad:
            tr = oracle.trace_path_logged(osc, cam, int(p), int(p) % n)
            e = PU.explain(PU.events_of(log[p]), tr, dev_cap=256)
            if shown < 12:
                ev = [(c, PU.draws(d, tr["d0"])) for c, d in PU.events_of(log[p])][:6]
                oe = [(int(c), PU.draws(int(d), tr["d0"])) for c, d in zip(tr["ev_code"], tr["ev_d"])][:6]
                print("  path", int(p), "px", int(p) % res, int(p) // res, "ours", got[p], "oracle", ora[p], "refgpu", r[p], e["kind"], e.get("margin"),
                      "dev", ev, "ora", oe)
                shown += 1
