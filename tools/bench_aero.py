#!/usr/bin/env python
"""Measurement of the aerodynamic twin on the geometry of bench.py's default workload:
`--bodies` unique 50k-face hulls resident in HBM, every face evaluated by sfh_aero_step_device, CUDA events on the launching
stream; then the same through sfh_aero_step with pinned host buffers (e2e) and a bounded CPU sample of the oracle's
orc_aero (fp32 sequential sums, one thread per body). Prints one JSON line. Not the headline bench: bench.py is.
  python tools/bench_aero.py [--bodies 1024] [--steps 20] [--warmup 3] [--wind uniform|jetpipe]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--wind", default="uniform", choices=["uniform", "jetpipe"])
    ap.add_argument("--cpu-bodies", type=int, default=32)
    args = ap.parse_args()

    import torch

    import bench
    from stonefish_b200 import BODY_STATE_DTYPE, INFO_DTYPE, HydroContext
    from stonefish_b200.hydro import Atmosphere, Jet, Pipe, Uniform

    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = HydroContext(0, stream=stream.cuda_stream)
    n = args.bodies
    keep = min(n, args.cpu_bodies)
    scene, states, host_meshes, faces, verts = bench.build_on_gpu(ctx, "cfg4", n, 0, keep)
    atm = Atmosphere(1.2)
    atm.AddVelocityField(Uniform((8.0, -3.0, 0.5)))
    if args.wind == "jetpipe":
        atm.AddVelocityField(Jet((0.0, -6.0, -1.0), (0.1, 1.0, 0.05), 0.5, 12.0))
        atm.AddVelocityField(Pipe((-8.0, 0.5, -1.5), (8.0, -0.5, -2.0), 2.0, 3.0, 4.0, 2.0))
    astruct = atm.to_struct()

    d_states = torch.empty(n * BODY_STATE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_infos = torch.empty(n * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_states.copy_(torch.from_numpy(states.view(np.uint8).reshape(-1).copy()))
    h_states = ctx.pinned_array(n, BODY_STATE_DTYPE)
    h_infos = ctx.pinned_array(n, INFO_DTYPE)
    h_states[:] = states
    torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            ctx.compute_aero_device(astruct, d_states.data_ptr(), n, d_infos.data_ptr())
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ctx.compute_aero_device(astruct, d_states.data_ptr(), n, d_infos.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        dev_ms = e0.elapsed_time(e1) / args.steps
        ctx.set_profiling(True)
        k_ms = []
        for _ in range(args.steps):
            ctx.compute_aero_device(astruct, d_states.data_ptr(), n, d_infos.data_ptr())
            k_ms.append(ctx.timings().submerged_ms)
        ctx.set_profiling(False)
        for _ in range(args.warmup):
            ctx.lib.sfh_aero_step(ctx.ctx, C.byref(astruct), h_states.ctypes.data, n, h_infos.ctypes.data)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.lib.sfh_aero_step(ctx.ctx, C.byref(astruct), h_states.ctypes.data, n, h_infos.ctypes.data)
        e2e_s = (time.perf_counter() - t0) / args.steps

    # parity of the sample against the oracle, then the CPU rate of the same sample
    from oracle import pyoracle
    from tests.parity import assert_counts_equal, max_rel_err

    gi = h_infos[:keep].copy()
    t0 = time.perf_counter()
    ri = pyoracle.aero(host_meshes, np.arange(keep), atm, states[:keep])
    cpu_s = time.perf_counter() - t0
    assert_counts_equal(gi, ri)
    err, _ = max_rel_err(gi, ri)
    cpu_faces = int(sum(m.n_faces for m in host_meshes[:keep]))

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    algo_bytes = 12 * faces + 12 * verts
    k = float(np.median(k_ms))
    print(json.dumps({
        "metric": "aero_triangle_evals_per_sec", "value": faces / (dev_ms * 1e-3), "unit": "triangle-evals/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
        "config": {"workload": f"aero-cfg4 ({args.wind} wind)", "bodies": n, "faces_per_step": faces, "l2": "inputs larger than L2"},
        "e2e": {"value": faces / e2e_s, "unit": "triangle-evals/s", "h2d_bytes_per_step": int(h_states.nbytes),
                "d2h_bytes_per_step": int(h_infos.nbytes)},
        "kernel_ms": k, "algorithmic_GBps": algo_bytes / (k * 1e-3) / 1e9, "peaks_file": peaks or None,
        "parity": {"bodies_checked": keep, "max_rel_err": err},
        "cpu_baseline": {"value": cpu_faces / cpu_s, "unit": "triangle-evals/s", "cores": 1, "kind": "port",
                         "sample": f"{keep} bodies ({cpu_faces} faces), orc_aero F64_ACC, one thread"}}))


if __name__ == "__main__":
    main()
