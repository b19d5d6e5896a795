"""Where does a path leave the oracle's path?  (test helper, CPU only)

The CUDA kernels and the CPU oracle draw the same XORWOW stream per path, so two
implementations agree on a path's radiance unless ONE compare somewhere along the path falls
the other way: a Woodcock accept `sigma_t / sigma_max < u'`, the exit test `t <= max_t`, the
roulette test or the Fresnel choice.  Device libm / fused arithmetic differ from the host's by
ulps, so such flips do happen (the < 1 % of paths the same-seed tests tolerate) -- but ONLY at
near-ties.  Both sides keep an event log per path -- (event code, draw counter at the event) --
and the oracle additionally logs every decision with the two values it compared
(oracle/cvr_oracle.h: cvro_trace).  `explain` finds the first event the logs disagree on,
identifies the decision that must have flipped, and returns how close to a tie the oracle saw
it.  A bug (wrong draw order, wrong lookup, stale table ...) shows up as disagreements at
decisions that were NOT close, or at places no decision explains.
"""
from __future__ import annotations

import numpy as np

D_STEP = 362437
D_INV = pow(D_STEP, -1, 2 ** 32)
EV_SCATTER, EV_BOUNDARY, EV_ESCAPE = 1, 2, 3
EVF_OK, EVF_WO_NEG, EVF_WI_NEG, EVF_KILLED, EVF_ZERO = 16, 32, 64, 128, 256
DEC_EXIT, DEC_ACCEPT, DEC_ROULETTE, DEC_FRESNEL = 1, 2, 3, 4


def draws(d: int, d0: int) -> int:
    """Number of draws consumed when the draw counter reads d (it started at d0)."""
    return ((int(d) - int(d0)) * D_INV) % 2 ** 32


def events_of(log) -> list[tuple[int, int]]:
    """(code, d) entries of a device log row (cap x 2 uint32; unused entries are zero)."""
    return [(int(c), int(d)) for c, d in log if c != 0]


def _decision_at(tr: dict, kind: int, d: int):
    m = np.nonzero((tr["dec_kind"] == kind) & (tr["dec_d"] == np.uint32(d)))[0]
    return None if len(m) == 0 else int(m[0])


def _first_decision_after(tr: dict, kind: int, d: int, d0: int):
    """first decision of `kind` whose draw comes after draw counter d"""
    n = draws(d, d0)
    for i in np.nonzero(tr["dec_kind"] == kind)[0]:
        if draws(tr["dec_d"][i], d0) > n:
            return int(i)
    return None


def explain(dev_events: list[tuple[int, int]], tr: dict, dev_cap: int = 0) -> dict:
    """Compare a device event list with the oracle trace of the same path.

    Returns {"kind": ..., "margin": float | None, "index": first differing event}:
      kind "same"        the event logs are identical (the radiances can then only differ by rounding)
           "accept"      a Woodcock accept test fell the other way; margin = |sigma_t/sigma_max - u'|
           "exit"        the segment-end test t <= max_t fell the other way; margin = |t - max_t| / max_t
           "roulette"    margin = |u - p_survive|
           "g1-zero"     one side's GGX weight is exactly 0 and the other's is not: GGX_G1 returns 0 when
                         1 - wo.z^2 <= 0 (GGX.h:232-236) and wo is not renormalised after refract, so a near-axial
                         refraction sits on a knife edge the last bit of wo.z decides (the reference's own host
                         and device builds disagree on these paths)
           "fresnel"     reflect / refract choice; margin = |u - F|
           "orientation" the GGX success / side flags differ but not through the Fresnel draw
           "truncated"   a log ran out of capacity before any difference
           "unexplained" anything else
    """
    d0 = tr["d0"]
    ora = list(zip((int(c) for c in tr["ev_code"]), (int(d) for d in tr["ev_d"])))
    n = min(len(dev_events), len(ora))
    i = next((k for k in range(n) if dev_events[k] != ora[k]), None)
    if i is None:
        if len(dev_events) == len(ora) and tr["n_events"] == len(ora):
            return {"kind": "same", "margin": None, "index": n}
        if (dev_cap and len(dev_events) >= dev_cap) or tr["n_events"] > len(ora):
            return {"kind": "truncated", "margin": None, "index": n}
        return {"kind": "unexplained", "margin": None, "index": n, "why": "one log is a strict prefix of the other"}
    (gc, gd), (oc, od) = dev_events[i], ora[i]
    out = {"index": i, "dev": (gc, draws(gd, d0)), "oracle": (oc, draws(od, d0))}
    if gd == od:
        if (gc & 0xF) != (oc & 0xF):
            # same draw count, different event: only escape-vs-hit can do that (the box test draws nothing)
            out.update(kind="unexplained" if EV_ESCAPE not in (gc & 0xF, oc & 0xF) else "isect", margin=None)
            return out
        diff = gc ^ oc
        if diff & EVF_ZERO:
            out.update(kind="g1-zero", margin=None)
            return out
        if diff & (EVF_OK | EVF_WO_NEG | EVF_WI_NEG):
            k = _first_decision_after(tr, DEC_FRESNEL, od, d0)
            if k is not None and draws(tr["dec_d"][k], d0) <= draws(od, d0) + 3 and not (diff & EVF_WI_NEG):
                out.update(kind="fresnel", margin=abs(float(tr["dec_a"][k]) - float(tr["dec_b"][k])))
            else:
                out.update(kind="orientation", margin=None)
            return out
        if diff & EVF_KILLED:
            k = _first_decision_after(tr, DEC_ROULETTE, od, d0)
            if k is None:
                out.update(kind="unexplained", margin=None, why="no roulette decision after the event")
            else:
                out.update(kind="roulette", margin=abs(float(tr["dec_a"][k]) - float(tr["dec_b"][k])))
            return out
        out.update(kind="unexplained", margin=None)
        return out
    # different draw counts: the side that stopped with FEWER draws decided to end the segment
    # where the other went on.  Its event tells which test that was.
    stop_code, stop_d = (gc, gd) if draws(gd, d0) < draws(od, d0) else (oc, od)
    base = stop_code & 0xF
    if base == EV_SCATTER:
        k = _decision_at(tr, DEC_ACCEPT, stop_d)
        if k is None:
            out.update(kind="unexplained", margin=None, why="the oracle drew no accept uniform at that draw")
        else:
            out.update(kind="accept", margin=abs(float(tr["dec_a"][k]) - float(tr["dec_b"][k])))
    elif base == EV_BOUNDARY:
        k = _decision_at(tr, DEC_EXIT, stop_d)
        if k is None:
            out.update(kind="unexplained", margin=None, why="the oracle made no exit test at that draw")
        else:
            a, b = float(tr["dec_a"][k]), float(tr["dec_b"][k])
            out.update(kind="exit", margin=abs(a - b) / max(abs(b), 1e-30))
    else:
        out.update(kind="isect", margin=None)
    return out


def summarise(results: list[dict]) -> dict:
    kinds: dict[str, int] = {}
    for r in results:
        kinds[r["kind"]] = kinds.get(r["kind"], 0) + 1
    margins = {k: sorted(r["margin"] for r in results if r["kind"] == k and r["margin"] is not None)
               for k in ("accept", "exit", "roulette", "fresnel")}
    return {"n": len(results), "kinds": kinds,
            "max_margin": {k: (v[-1] if v else None) for k, v in margins.items()},
            "median_margin": {k: (v[len(v) // 2] if v else None) for k, v in margins.items()}}
