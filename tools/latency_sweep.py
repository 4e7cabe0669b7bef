#!/usr/bin/env python
"""Host-to-host step latency of single hulls of 1k .. 900k faces (submerged, on a flat sea, crossing waves) and of the
FluidDynamicsTest scene, through the Python wrapper of the tree given as argument (default: this one). Prints one line, so that two
trees or two library builds (SFH_LIB=...) can be compared on the same box within one gpurun call:
    python tools/latency_sweep.py            # this tree
    python tools/latency_sweep.py /path/to/another/checkout
The same-box A/B rows of DESIGN.md 7b (direct mode, global wave bound, two waves of warps, round 1 against round 2) come from it."""
import os, sys, time
ROOT = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stonefish_b200 import HydroContext, scenes


def time_gpu(ctx, sc, steps=400):
    st = sc.bind(ctx)
    for _ in range(30):
        ctx.compute(sc.ocean, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.compute(sc.ocean, st)
    return (time.perf_counter() - t0) / steps * 1e6


ctx = HydroContext(0)
cases = [("cfg1", scenes.config1())]
for seg, stk in ((24, 22), (50, 51), (100, 101), (400, 251), (560, 358), (600, 380), (800, 500), (900, 500)):
    n = 2 * seg * (stk - 1)
    cases.append((f"sub{n}", scenes.config2(seg, stk)))
    cases.append((f"flat{n}", scenes.config3(seg, stk, False)))
    cases.append((f"cross{n}", scenes.config3(seg, stk, True)))
print(ROOT, " ".join(f"{n}={time_gpu(ctx, sc):.1f}" for n, sc in cases))
