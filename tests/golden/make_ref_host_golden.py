"""Generate tests/golden/ref_host_*.npz from the REFERENCE's own host-compiled
functions (oracle/_ref/libcvr_ref_host.so, built by oracle/Makefile from the headers
under /root/reference).  Run in the build container:  python tests/golden/make_ref_host_golden.py

The vectors pin oracle/cvr_oracle.c on boxes where /root/reference is absent.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def f(a):
    return np.ascontiguousarray(a, np.float32)


def gen_inputs(seed=1234):
    rng = np.random.default_rng(seed)
    n = 4096
    wi = rng.normal(size=(n, 3))
    wi /= np.linalg.norm(wi, axis=1, keepdims=True)
    wi = f(wi)
    wi[::50] = f([0, 0, 1])
    wi[25::50] = f([0, 0, -1])
    wi[7::77, 2] *= np.float32(1e-3)
    wi[13::301, 2] = 0.0
    u = f(rng.random((n, 3)))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = f(d)
    d[5::11, 0] = 0.0
    o = f(rng.uniform(-1.5, 1.5, (n, 3)))
    o[::3] = f(rng.uniform(-0.5, 0.5, (len(o[::3]), 3)))
    o[4::7, 1] = np.float32(0.5)
    o[9::14, 2] = np.float32(-0.5)
    g = f(np.where(np.arange(n) % 2 == 0, 0.0, rng.uniform(-0.9, 0.9, n)))
    return dict(wi=wi, u=u, d=d, o=o, g=g)


def main():
    R = B.ref_host()
    if R is None:
        raise SystemExit("oracle/_ref/libcvr_ref_host.so not built (needs /root/reference)")
    inp = gen_inputs()
    n = len(inp["wi"])
    alpha = f([0.1, 0.1])
    eta = C.c_float()
    a2 = f([0, 0])
    R.ref_ggx_defaults(B.fp(a2), C.byref(eta))
    ggx_ok = np.zeros(n, np.int32)
    ggx_wo = np.zeros((n, 3), np.float32)
    ggx_w = np.zeros(n, np.float32)
    ggx_used = np.zeros(n, np.int32)
    hg = np.zeros((n, 3), np.float32)
    ab_hit = np.zeros(n, np.int32)
    ab_dist = np.zeros(n, np.float32)
    ab_n = np.zeros((n, 3), np.float32)
    ab_in = np.zeros(n, np.int32)
    fr_f = np.zeros(n, np.float32)
    fr_ct = np.zeros(n, np.float32)
    g1 = np.zeros(n, np.float32)
    mort = np.zeros(n, np.uint32)
    bmin, bmax = f([-0.5, -0.5, -0.5]), f([0.5, 0.5, 0.5])
    for i in range(n):
        wo = f([9, 9, 9])
        w = C.c_float()
        k = C.c_int()
        ggx_ok[i] = R.ref_ggx_sample(B.fp(alpha), eta, B.fp(inp["wi"][i]), B.fp(inp["u"][i]),
                                     B.fp(wo), C.byref(w), C.byref(k))
        ggx_wo[i], ggx_w[i], ggx_used[i] = wo, w.value, k.value
        out = f([0, 0, 0])
        R.ref_hg_sample(B.fp(inp["d"][i]), inp["g"][i], inp["u"][i][0], inp["u"][i][1], B.fp(out))
        hg[i] = out
        nn = f([0, 0, 1])
        dist = C.c_float()
        ins = C.c_int()
        ab_hit[i] = R.ref_aabb_intersect(B.fp(bmin), B.fp(bmax), B.fp(inp["o"][i]), B.fp(inp["d"][i]),
                                         C.byref(dist), B.fp(nn), C.byref(ins))
        ab_dist[i], ab_n[i], ab_in[i] = dist.value, nn, ins.value
        ct = C.c_float()
        fr_f[i] = R.ref_fresnel_dielectric(eta, inp["wi"][i][2], C.byref(ct))
        fr_ct[i] = ct.value
        g1[i] = R.ref_ggx_g1(B.fp(alpha), B.fp(inp["d"][i]), B.fp(inp["wi"][i]))
        mort[i] = R.ref_morton3d(*[float(x) for x in inp["o"][i]])
    frames = []
    for nrm in [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]:
        xs = [f([0] * 3) for _ in range(3)]
        R.ref_frame_from_z(B.fp(f(nrm)), *[B.fp(x) for x in xs])
        frames.append(np.stack(xs))
    np.savez_compressed(
        os.path.join(HERE, "ref_host_golden.npz"), eta=np.float32(eta.value), alpha=alpha,
        ggx_ok=ggx_ok, ggx_wo=ggx_wo, ggx_w=ggx_w, ggx_used=ggx_used, hg=hg, ab_hit=ab_hit,
        ab_dist=ab_dist, ab_n=ab_n, ab_in=ab_in, fr_f=fr_f, fr_ct=fr_ct, g1=g1, mort=mort,
        frames=np.stack(frames), **{"in_" + k: v for k, v in inp.items()})
    print("wrote ref_host_golden.npz;  ggx failures:", int((ggx_ok == 0).sum()), "of", n)


if __name__ == "__main__":
    main()
