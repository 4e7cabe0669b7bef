#!/usr/bin/env python
"""bench.py — hydrodynamic triangle-evals/s of the B200 path (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg4-shared|cfg2|cfg3|cfg3-flat|cfg5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's CPU algorithm (oracle port) on the host cores

A "step" = one recompute of the hydrodynamic forces of every body of the workload (one pass of the hot path).
Default workload `cfg4` = BASELINE.json configs[3], the config the metric's 1/2/4/8-GPU figures are quoted on:
4096 independent 50k-triangle AUV instances PER GPU (weak scaling), each with its own mesh (3.3 GB per GPU, far
larger than the 126 MB L2, so nothing is cache resident between steps and no L2 flush is needed).
  value    : whole-job triangle-evals/s with meshes AND body states resident in HBM (sfh_step_device), device-timed
             with CUDA events on the launching stream, max over ranks.
  e2e      : same metric through the C ABI call a Stonefish build would make (sfh_step) with HOST buffers: the
             pinned H2D copy of every body state and the D2H copy of every result are inside the timed region.
  roofline : the per-face kernel alone (CUDA events inside libsfh around that launch), algorithmic bytes
             (12 B/face + 12 B/vertex) / its mean duration / measured HBM peak.
  cpu_baseline : the oracle in the reference's own mode (fp32 sequential sums, one task per body on all host
             cores) on a bounded sample of the same instances, rank 0 at N=1 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hydrodynamic triangle-evals/sec"
UNIT = "triangle-evals/s"
ALGO_BYTES_PER_FACE_NOTE = "12 B/face (3 x u32 ids) + 12 B/unique vertex (3 x f32)"


# ------------------------------------------------------------------------------------------------ workloads
def workload_spec(name: str):
    from stonefish_b200 import scenes

    specs = {
        # name: (scene factory(n, first) -> Scene with ONE shared base mesh, bodies per GPU, unique meshes, description)
        "cfg4": (lambda n, first: scenes.config4(n, 250, 101, False, first), 4096, True,
                 "4096 independent 50k-triangle AUV instances per GPU, unique mesh per instance, fully submerged, uniform current (BASELINE configs[3])"),
        "cfg4-shared": (lambda n, first: scenes.config4(n, 250, 101, False, first), 4096, False,
                        "as cfg4 but all instances share one mesh (L2 resident; not a roofline figure)"),
        "cfg2": (lambda n, first: scenes.config2(), 1, False, "single 200k-triangle hull, submerged, uniform current (BASELINE configs[1])"),
        "cfg3": (lambda n, first: scenes.config3(400, 251, True), 1, False, "single 200k-triangle vessel crossing a wavy surface (BASELINE configs[2])"),
        "cfg3-flat": (lambda n, first: scenes.config3(400, 251, False), 1, False, "single 200k-triangle vessel crossing a flat surface"),
        "cfg5": (lambda n, first: scenes.config5(n, 1000, 501, False, first), 128, True,
                 "128 vessels per GPU with 1M-triangle hulls, unique meshes, waves, Jet+Pipe currents (BASELINE configs[4])"),
    }
    return specs[name]


def jitter_seed(workload: str, instance: int) -> int:
    return (0xC0FFEE if workload == "cfg5" else 0xB200) + instance


def build_on_gpu(ctx, workload: str, n_bodies: int, first: int, keep_host: int):
    """Scene states for instances first..first+n-1 and their meshes uploaded to ctx.
    Returns (scene, states with device mesh ids, host meshes of the first `keep_host` bodies, faces, verts)."""
    from stonefish_b200 import meshes as M

    factory, _, unique, _ = workload_spec(workload)
    scene = factory(n_bodies, first)
    base = scene.meshes[0]
    states = scene.states.copy()
    host_meshes = []
    if unique:
        def make(i):
            return M.jitter(base, jitter_seed(workload, first + i))

        def up(i):
            m = make(i)
            return ctx.upload_mesh(m), (m if i < keep_host else None)

        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            res = list(pool.map(up, range(n_bodies)))
        states["mesh_id"] = np.array([r[0] for r in res], dtype=np.uint32)
        host_meshes = [r[1] for r in res[:keep_host]]
    else:
        mid = ctx.upload_mesh(base)
        states["mesh_id"] = mid
        host_meshes = [base] * min(keep_host, n_bodies)
    faces = base.n_faces * n_bodies
    verts = base.n_verts * n_bodies
    return scene, states, host_meshes, faces, verts


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi samples DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thread, self.idx, self.first = [], None, None, gpu_index, 0

    def mark(self):
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in (self.rows[self.first:] or self.rows):  # a very short timed region: fall back to warm-up samples
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def cpu_stepper(meshes, body_mesh, ocean, states):
    """The CPU arm: (step(n_threads) -> (infos, seconds), kind). "reference" = the reference's OWN definitions of the path compiled
    from its source text (oracle/_ref/librefsrc.so, built where the reference checkout is mounted and shipped with the repo
    snapshot; real Bullet, a stand-in for the GLM primitives — oracle/ref_harness.cpp); "port" = the oracle's restatement, bit-identical
    to it (tests/test_oracle_refsource.py), used when that library is absent."""
    from oracle import pyoracle as po
    from oracle import pyrefsrc as rs

    if rs.available():
        sc = rs.RefScene(meshes, body_mesh, ocean, states)
        return (lambda threads: sc.step(threads)), "reference"
    osc = po.OracleScene(meshes, body_mesh, ocean, states)
    return (lambda threads: osc.step(po.ORC_REF_F32_SEQ, n_threads=threads)), "port"


def cpu_oracle_rate(workload: str, scene, host_meshes, budget_s: float, threads: int):
    """The reference's CPU path on the first len(host_meshes) bodies, scheduled like the reference: one task per body on `threads`
    pool threads; single-body workloads are therefore single-threaded in the reference's schedule.
    Returns (faces/s, sample description, faces per pass, seconds per pass, kind)."""
    n = len(host_meshes)
    step, kind = cpu_stepper(host_meshes, np.arange(n), scene.ocean, scene.states[:n])
    faces = sum(m.n_faces for m in host_meshes)
    secs, passes = 0.0, 0
    t_end = time.perf_counter() + budget_s
    while passes < 1 or (time.perf_counter() < t_end and passes < 1000):
        _, s = step(threads)
        secs += s
        passes += 1
    what = "the reference's own code (oracle/_ref/librefsrc.so)" if kind == "reference" else "oracle port"
    desc = f"{n} bodies ({faces} faces) of the workload x {passes} passes, one task per body on {threads} threads, fp32 sequential sums, {what}"
    return faces * passes / secs, desc, faces, secs / passes, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path — its definitions compiled from its source text
    (oracle/_ref/librefsrc.so; the full reference build needs glm/SDL2/GL, absent here), else the bit-identical oracle port —
    on all host threads, same config/metric/unit; a step = one pass over a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    from stonefish_b200 import meshes as M

    po.build()
    factory, per_gpu, unique, desc = workload_spec(args.workload)
    threads = os.cpu_count() or 1
    sample_bodies = min(per_gpu, 4 * threads if per_gpu > 1 else 1)
    scene = factory(sample_bodies, 0)
    base = scene.meshes[0]
    meshes = [M.jitter(base, jitter_seed(args.workload, i)) for i in range(sample_bodies)] if unique else [base] * sample_bodies
    step, kind = cpu_stepper(meshes, np.arange(sample_bodies), scene.ocean, scene.states)
    faces = sum(m.n_faces for m in meshes)
    use_threads = threads if sample_bodies > 1 else 1
    for _ in range(args.warmup):
        step(use_threads)
    t = 0.0
    for _ in range(args.steps):
        _, s = step(use_threads)
        t += s
    value = faces * args.steps / t
    what = "the reference's own code compiled from its source text (oracle/_ref/librefsrc.so)" if kind == "reference" else "oracle port of the reference"
    sample = f"{sample_bodies} bodies ({faces} faces) per step, one task per body on {use_threads} threads (reference schedule), fp32 sequential sums, {what}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "bodies_per_step": sample_bodies, "faces_per_step": faces},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": use_threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from stonefish_b200 import BODY_STATE_DTYPE, INFO_DTYPE, HydroContext
    from stonefish_b200.sharding import checksum, gather_infos

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run with --nproc-per-node {args.gpus}")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    factory, per_gpu, unique, desc = workload_spec(args.workload)
    n = args.bodies if args.bodies else per_gpu
    stream = torch.cuda.Stream(device=dev)
    ctx = HydroContext(local_rank, stream=stream.cuda_stream)
    keep = min(n, 4 * (os.cpu_count() or 1)) if (rank == 0 and world == 1 and not args.no_cpu) else 0
    t_setup = time.perf_counter()
    scene, states, host_meshes, faces, verts = build_on_gpu(ctx, args.workload, n, rank * n, keep)
    t_setup = time.perf_counter() - t_setup
    ocean = scene.ocean
    ctx._sync_waves(ocean)
    ostruct = ocean.to_struct()
    algo_bytes = 12 * faces + 12 * verts + (256 + 208) * n + (1048576 if ocean.hasWaves() else 0)

    # device-resident state/result buffers (value) and pinned host buffers (e2e)
    d_states = torch.empty(n * BODY_STATE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_infos = torch.empty(n * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_states.copy_(torch.from_numpy(states.view(np.uint8).reshape(-1).copy()))
    h_states = ctx.pinned_array(n, BODY_STATE_DTYPE)
    h_infos = ctx.pinned_array(n, INFO_DTYPE)
    h_states[:] = states
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def step_device():
        ctx.compute_device(ostruct, d_states.data_ptr(), n, d_infos.data_ptr())

    def step_host():
        ctx.lib.sfh_step(ctx.ctx, C.byref(ostruct), h_states.ctypes.data, n, h_infos.ctypes.data)

    # ---- value: device-resident, CUDA events on the launching stream
    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank)
        sampler.start()  # nvidia-smi needs ~0.1-0.3 s to start streaming: launch it before the warm-up
        for _ in range(args.warmup):
            step_device()
        barrier()
        sampler.mark()   # samples from here to stop() are inside the timed regions (device, kernel and e2e loops)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)

        # ---- roofline: the per-face kernel alone, timed by events inside libsfh around that launch
        ctx.set_profiling(True)
        k_ms, pro_ms, epi_ms, launches = [], [], [], 0
        for _ in range(args.steps):
            step_device()
            t = ctx.timings()
            k_ms.append(max(t.submerged_ms, t.surface_ms))
            pro_ms.append(t.prologue_ms); epi_ms.append(t.epilogue_ms)
            launches = t.launches
            dominant = "k_submerged" if t.submerged_ms >= t.surface_ms else "k_surface"
        ctx.set_profiling(False)
        barrier()

        # ---- e2e: host buffers through sfh_step, H2D + D2H inside the timed region (wall clock between barriers)
        for _ in range(args.warmup):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop()

    # max over ranks
    times = torch.tensor([dev_ms, e2e_s * 1e3, float(np.mean(k_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kern_ms = (float(x) for x in times.cpu())

    # ---- validation gather (C1), outside every timed region: per-body results + checksum over NCCL
    infos = h_infos.copy()
    if world > 1:
        all_infos, sums = gather_infos(infos, device=dev)
    else:
        all_infos, sums = infos, [checksum(infos) & 0x7FFFFFFFFFFFFFFF]
    ok = bool(np.isfinite(all_infos["Fdq"]).all() and (all_infos["n_faces"] > 0).all())

    if rank == 0:
        total_faces = faces * world
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of the same
        # kernel on the same meshes (profiles/ncu_traffic_r01.json: dram__bytes_read.sum + dram__bytes_write.sum at
        # 1024 bodies), scaled by face count; null for workloads that capture does not describe.
        traffic, traffic_src, secondary = None, None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")
        if args.workload == "cfg4" and dominant == "k_submerged" and os.path.exists(tpath):
            t = json.load(open(tpath))
            traffic = int(t["dram_bytes_per_face"] * faces)
            traffic_src = f"{t['dram_bytes_per_face']:.2f} B/face measured by ncu ({t['report']}) x {faces} faces"
            # secondary ceilings of the same capture: the kernel is fp32-issue-bound, not HBM-bound
            secondary = dict(t.get("secondary") or {}, source=t["report"])
        line = {
            "metric": METRIC, "value": total_faces * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 per-face terms, f64 accumulation", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "bodies_per_gpu": n, "faces_per_gpu_per_step": faces,
                       "unique_meshes": unique, "l2": "inputs larger than L2 (no flush)" if algo_bytes > 4 * 126e6 else "inputs NOT larger than L2; steps re-read a cache-resident mesh",
                       "setup_s": round(t_setup, 2)},
            "e2e": {"value": total_faces * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(n * BODY_STATE_DTYPE.itemsize),  # body states; the ocean record (<= 6.2 KB) is only re-sent when it changes "d2h_bytes_per_step": int(n * INFO_DTYPE.itemsize + 4),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes),
                         "kernel_ms": kern_ms, "prologue_ms": float(np.mean(pro_ms)), "epilogue_ms": float(np.mean(epi_ms)),
                         "bytes_per_face": ALGO_BYTES_PER_FACE_NOTE, "secondary": secondary},
            "clocks": clocks,
            "validation": {"checksums": [f"{s:016x}" for s in sums], "bodies_gathered": int(all_infos.shape[0]), "finite": ok},
        }
        if keep:
            from oracle import pyoracle as po

            po.build()
            threads = os.cpu_count() or 1
            use = threads if keep > 1 else 1
            rate, sample, _, _, kind = cpu_oracle_rate(args.workload, scene, host_meshes, args.cpu_seconds, use)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": use, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg4-shared", "cfg2", "cfg3", "cfg3-flat", "cfg5"])
    ap.add_argument("--bodies", type=int, default=0, help="bodies per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU time budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
