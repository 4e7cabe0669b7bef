#!/usr/bin/env python
"""Measurement of the device wave-field producer: time per sfh_wavegen_simulate (h(k,t) pass + 16 FFT
passes over 3 layers of 256^2 RGBA32F + the planar copy), CUDA events on the launching stream, against the oracle's CPU
restatement of the same passes; plus the read-back the reference needs every frame and this path does not.
  python tools/bench_wavegen.py [--steps 200] [--passes 8]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--passes", type=int, default=8)
    args = ap.parse_args()
    import torch

    from oracle import pyoracle
    from stonefish_b200 import HydroContext
    from stonefish_b200.hydro import wave_params

    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = HydroContext(0, stream=stream.cuda_stream)
    p = wave_params(2.0)
    p.passes = args.passes
    N = 1 << args.passes
    t0 = time.perf_counter()
    ctx.wavegen_init(p)
    init_s = time.perf_counter() - t0
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            ctx.wavegen_simulate(0.016 * i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            ctx.wavegen_simulate(0.016 * i)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        dev_us = e0.elapsed_time(e1) * 1e3 / args.steps
        t0 = time.perf_counter()
        for i in range(20):
            ctx.wavegen_simulate(0.016 * i)
            ctx.wavegen_download()
        dl_us = (time.perf_counter() - t0) * 1e6 / 20
    got = ctx.wavegen_download().astype(np.float64)
    w = pyoracle.WaveOracle(p)
    t0 = time.perf_counter()
    ref = w.simulate(0.016 * 19).astype(np.float64)
    cpu_us = (time.perf_counter() - t0) * 1e6
    err = max(np.abs(got[l] - ref[l]).max() / np.abs(ref[l]).max() for l in range(3))
    # bytes one simulate moves at minimum: spectrum in (2 x N^2 x 16), 3 layers out + 3 in + 3 out (tmp round trip), planes out
    algo_bytes = N * N * 16 * (2 + 3 + 3 + 3) + 2 * 4 * (N * N + N + 1)
    print(json.dumps({"metric": "wave_field_updates_per_sec", "value": 1e6 / dev_us, "unit": "fields/s", "us_per_simulate": dev_us,
                      "fft_size": N, "layers": 3, "launches_per_simulate": 3, "steps": args.steps,
                      "us_per_simulate_plus_download": dl_us, "host_init_ms": init_s * 1e3,
                      "algorithmic_GBps": algo_bytes / (dev_us * 1e-6) / 1e9,
                      "cpu_baseline": {"us_per_simulate": cpu_us, "kind": "port", "cores": 1, "sample": "one orc_wave_simulate call"},
                      "parity": {"max_rel_err_vs_oracle": err, "t": 0.016 * 19}}))


if __name__ == "__main__":
    main()
