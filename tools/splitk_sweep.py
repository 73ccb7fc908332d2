#!/usr/bin/env python
"""Split-K sweep of tc_pixgemm at the bench shapes: device time per launch (tools/tc_sweep.py::one: 20 back-to-back launches,
warm L2) for the unsplit plan, the cycle model's choice and forced (tile, S) pairs, in the two in-step forms of a launch
(bf16 output; fp32 pre-norm output + fused statistics).  GG_PROF=1 adds the per-CTA clock64 breakdown."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tc_sweep  # noqa: E402

CONFIGS = [("unsplit", dict(GG_TC_SPLITK="1")), ("auto", dict(GG_TC_SPLITK="auto")),
           ("bn256 S8", dict(GG_TC_SPLITK="8", GG_TC_BN="256")), ("bn256 S4", dict(GG_TC_SPLITK="4", GG_TC_BN="256")),
           ("bn256 S2", dict(GG_TC_SPLITK="2", GG_TC_BN="256")),
           ("bn128 S4", dict(GG_TC_SPLITK="4", GG_TC_BN="128")), ("bn128 S2", dict(GG_TC_SPLITK="2", GG_TC_BN="128"))]

if __name__ == "__main__":
    shapes = sys.argv[1:] or ["g_h1", "d_h3", "g_h2", "d_h2"]
    for form in ("bf16", "stats"):
        os.environ.pop("GG_SWEEP_STATS", None)
        if form == "stats":
            os.environ["GG_SWEEP_STATS"] = "1"
        for shape in shapes:
            for op in ("down", "up"):
                for name, env in CONFIGS:
                    for k in ("GG_TC_SPLITK", "GG_TC_BN"):
                        os.environ.pop(k, None)
                    os.environ.update(env)
                    print("CONFIG form=%s %s" % (form, name), flush=True)
                    try:
                        tc_sweep.one(shape, op)
                    except Exception as e:       # a forced pair the layer cannot take
                        print("SKIP", shape, op, name, str(e)[:120], flush=True)
