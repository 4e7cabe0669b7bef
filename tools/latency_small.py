#!/usr/bin/env python
"""Step latency of small scenes through the public host API (sfh_step with pageable host buffers, CUDA-graph replay): the
FluidDynamicsTest scene (25 bodies, 23k faces) and single bodies of 1k..200k faces, next to the oracle port on one thread and
on one thread per body. Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle
from stonefish_b200 import HydroContext, scenes


def time_gpu(ctx, sc, steps=300):
    st = sc.bind(ctx)
    for _ in range(20):
        ctx.compute(sc.ocean, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.compute(sc.ocean, st)
    return (time.perf_counter() - t0) / steps * 1e6


def time_cpu(sc, threads):
    o = pyoracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states)
    o.step(0, n_threads=threads)
    secs = [o.step(0, n_threads=threads)[1] for _ in range(5)]
    return float(np.median(secs)) * 1e6


ctx = HydroContext(0)
rows = []
cases = [("FluidDynamicsTest scene", scenes.config1())]
for seg, stk in ((24, 22), (50, 51), (100, 101), (400, 251)):
    cases.append((f"single hull {2 * seg * (stk - 1)} faces, submerged", scenes.config2(seg, stk)))
    cases.append((f"single hull {2 * seg * (stk - 1)} faces, crossing waves", scenes.config3(seg, stk, True)))
for name, sc in cases:
    rows.append({"scene": name, "bodies": sc.n_bodies, "faces": sc.faces_per_step(), "gpu_us_per_step": round(time_gpu(ctx, sc), 1),
                 "cpu_1_thread_us": round(time_cpu(sc, 1), 1), "cpu_thread_per_body_us": round(time_cpu(sc, min(16, sc.n_bodies)), 1)})
print(json.dumps({"metric": "host-to-host step latency", "unit": "us", "rows": rows}))
