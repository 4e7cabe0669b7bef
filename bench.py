#!/usr/bin/env python
"""bench.py — hydrodynamic triangle-evals/s of the B200 path (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg4-shared|cfg2|cfg3|cfg3-flat|cfg5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's own CPU code of the path on the host cores

A "step" = one recompute of the hydrodynamic forces of every body of the workload (one pass of the hot path).
Headline workload `cfg4` = BASELINE.json configs[3], the config the metric's 1/2/4/8-GPU figures are quoted on:
4096 independent 50k-triangle AUV instances PER GPU (weak scaling), each with its own mesh (3.3 GB per GPU, far
larger than the 126 MB L2, so nothing is cache resident between steps and no L2 flush is needed).
  value    : whole-job triangle-evals/s with meshes AND body states resident in HBM (sfh_step_device), device-timed
             with CUDA events on the launching stream, max over ranks.
  e2e      : same metric through the C ABI call a Stonefish build would make (sfh_step) with HOST buffers: the
             pinned H2D copy of every body state and the D2H copy of every result are inside the timed region.
  roofline : the per-face kernel(s) alone (CUDA events inside libsfh around those launches), algorithmic bytes
             (12 B/face + 12 B/vertex) / mean duration / measured HBM peak.
  validation : after the timed loops the results of the first bodies — produced by the very kernels that were timed
             (DEBUG=0 templates, full-size meshes) — are compared with the oracle in F64_ACC mode: counters equal,
             sums within 1e-9 of the force norm. A failure makes the run exit non-zero.
  workloads: with the default invocation the other BASELINE configs (cfg2, cfg3, cfg3-flat, cfg5) are measured in the
             same process after the headline, each to the same rules (value, e2e, roofline, validation).
  strong   : under torchrun (N > 1) BASELINE configs[3] AS WRITTEN — 4096 instances in total, block-partitioned over the
             N ranks — next to the weak-scaling headline, checked against rank 0's own results for the same instances.
  cpu_baseline : the reference's own CPU code (oracle/_ref/librefsrc.so, else the oracle port) in the reference's schedule
             (fp32 sequential sums, one task per body on all host cores) on a bounded sample, rank 0 at N=1 only.
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
REL_TOL = 1e-9  # north_star: per-body sums within 1e-9 relative to the force norm
VALIDATE_BODIES, VALIDATE_FACES = 64, 36_000_000  # oracle check after the timed loops: first bodies, bounded face count
EXTRA_WORKLOADS_1GPU = ("cfg2", "cfg3", "cfg3-flat", "cfg5")
EXTRA_WORKLOADS_NGPU = ("cfg5",)  # 128 vessels per GPU: at N = 8 this is BASELINE configs[4] as written (1024 vessels)


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


def config_dict(workload: str, n: int) -> dict:
    """The `config` object: a function of the workload and its size only, so both arms (ours / --impl reference) print the same one."""
    factory, _, unique, desc = workload_spec(workload)
    base = factory(1, 0).meshes[0]
    faces, verts = base.n_faces * n, base.n_verts * n
    algo = 12 * faces + 12 * verts
    return {"workload": workload, "description": desc, "bodies_per_gpu": n, "faces_per_gpu_per_step": faces, "unique_meshes": unique,
            "l2": "inputs larger than L2 (no flush)" if algo > 4 * 126e6 else "inputs NOT larger than L2; steps re-read a cache-resident mesh"}


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
        def up(i):
            m = M.jitter(base, jitter_seed(workload, first + i))
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
            self.rows.append([time.perf_counter()] + [t.strip() for t in line.split(",")])

    def stop(self, trace: bool = False) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, ts, reasons = [], [], [], [], set()
        for r in (self.rows[self.first:] or self.rows):  # a very short timed region: fall back to warm-up samples
            if len(r) < 10:
                continue
            try:
                sm.append(float(r[2])); mx.append(float(r[3])); pw.append(float(r[4])); ts.append(r[0])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if sm:
            out["sm_mhz_min"], out["power_w_max"] = min(sm), max(pw)
        if trace and sm:
            t0 = ts[0]
            out["trace"] = [[round(t - t0, 3), s, p] for t, s, p in zip(ts, sm, pw)]
        return out


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


def validate_against_oracle(scene, host_meshes, infos: np.ndarray) -> dict:
    """The timed kernels' results for the first bodies against the oracle in ORC_F64_ACC mode (identical fp32 per-face terms, fp64
    sums: the contract of the CUDA path). Counters must be equal, sums within REL_TOL of the force norm."""
    from oracle import pyoracle as po
    from tests.parity import COUNT_FIELDS, max_rel_err

    k, faces = 0, 0
    while k < len(host_meshes) and k < VALIDATE_BODIES and (k == 0 or faces + host_meshes[k].n_faces <= VALIDATE_FACES):
        faces += host_meshes[k].n_faces
        k += 1
    osc = po.OracleScene(host_meshes[:k], np.arange(k), scene.ocean, scene.states[:k])
    ref, _ = osc.step(po.ORC_F64_ACC, n_threads=min(k, os.cpu_count() or 1))
    got = infos[:k]
    err, where = max_rel_err(got, ref)
    bad = [f for f in COUNT_FIELDS if not np.array_equal(got[f], ref[f])]
    # Jet / Pipe: CUDA's exp / pow differ from glibc's in the last bit for ~1e-4 of the samples, which may flip a
    # drag-branch flag with probability ~1e-8 per face; the classification (class counts, degenerate) never depends on currents
    flips = int(sum(abs(int(a) - int(b)) for f in ("n_quadratic", "n_skin") for a, b in zip(got[f], ref[f])))
    general = any(c.isEnabled() and type(c).__name__ != "Uniform" for c in scene.ocean.currents) and scene.ocean.currentsEnabled
    hard_bad = [f for f in bad if f not in ("n_quadratic", "n_skin")] if general else bad
    counts_equal = not hard_bad and (flips <= max(1, faces // 10_000_000) if general else flips == 0)
    return {"against": "oracle ORC_F64_ACC (oracle/hydro_oracle.cpp)", "bodies": k, "faces": int(faces), "max_rel_err": err,
            "worst": list(where) if where else None, "tolerance": REL_TOL, "counts_equal": bool(counts_equal),
            "count_fields_differing": bad, "drag_flag_flips": flips, "ok": bool(err <= REL_TOL and counts_equal)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path — its definitions compiled from its source text
    (oracle/_ref/librefsrc.so; the full reference build needs glm/SDL2/GL, absent here), else the bit-identical oracle port —
    on all host threads, same config/metric/unit as our arm; a step = one pass over ALL bodies of the per-GPU workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    from stonefish_b200 import meshes as M

    po.build()
    factory, per_gpu, unique, desc = workload_spec(args.workload)
    threads = os.cpu_count() or 1
    n = args.bodies if args.bodies else per_gpu
    scene = factory(n, 0)
    base = scene.meshes[0]
    if unique:
        with ThreadPoolExecutor(max_workers=min(16, threads)) as pool:
            meshes = list(pool.map(lambda i: M.jitter(base, jitter_seed(args.workload, i)), range(n)))
    else:
        meshes = [base] * n
    step, kind = cpu_stepper(meshes, np.arange(n), scene.ocean, scene.states)
    faces = sum(m.n_faces for m in meshes)
    use_threads = threads if n > 1 else 1
    for _ in range(args.warmup):
        step(use_threads)
    t = 0.0
    for _ in range(args.steps):
        _, s = step(use_threads)
        t += s
    value = faces * args.steps / t
    what = "the reference's own code compiled from its source text (oracle/_ref/librefsrc.so)" if kind == "reference" else "oracle port of the reference"
    sample = f"all {n} bodies ({faces} faces) per step, one task per body on {use_threads} threads (reference schedule), fp32 sequential sums, {what}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, n),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": use_threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
class Rig:
    """Process-wide state of our arm: device, stream, ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            if self.world == 1 and args.gpus > 1:
                raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run with --nproc-per-node {args.gpus}")
            args.gpus = self.world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]


def measure(rig: Rig, args, workload: str, n: int, first: int, keep_host: int, sampler: ClockSampler | None = None) -> dict:
    """One workload on this rank's GPU: upload, warm up, time K device-resident steps (value), K profiled steps (per-kernel events)
    and K host-buffer steps (e2e). Times are the max over ranks. Returns the measurements and the last step's results."""
    torch = rig.torch
    from stonefish_b200 import BODY_STATE_DTYPE, INFO_DTYPE, HydroContext

    dev, stream = rig.dev, rig.stream
    ctx = HydroContext(rig.local_rank, stream=stream.cuda_stream)
    t_setup = time.perf_counter()
    scene, states, host_meshes, faces, verts = build_on_gpu(ctx, workload, n, first, keep_host)
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

    # what sfh_step derives from host-side states, the caller of a device-resident step says itself: which paths its bodies can take
    modes = set(int(m) for m in np.unique(states["mode"]))
    hints = (1 if modes <= {0, 1} else 0) | (2 if modes <= {0, 2} else 0)  # SFH_HINT_NO_SURFACE | SFH_HINT_NO_SUBMERGED

    def step_device():
        ctx.compute_device(ostruct, d_states.data_ptr(), n, d_infos.data_ptr(), hints)

    sfh_step, ocean_ref, p_states, p_infos = ctx.lib.sfh_step, C.byref(ostruct), h_states.ctypes.data, h_infos.ctypes.data

    def step_host():  # the C call a host makes per step: pointers and the ocean record are the caller's, prepared once
        if sfh_step(ctx.ctx, ocean_ref, p_states, n, p_infos) != 0:
            raise RuntimeError("sfh_step: " + ctx.lib.sfh_last_error(ctx.ctx).decode())

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        rig.barrier()
        if sampler is not None:
            sampler.mark()   # samples from here to stop() are inside the timed regions (device, kernel and e2e loops)
        # ---- value: device-resident, CUDA events on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        rig.barrier()
        dev_ms = e0.elapsed_time(e1)

        # ---- roofline: the per-face kernels alone, timed by events inside libsfh around those launches
        ctx.set_profiling(True)
        k_ms, pro_ms, epi_ms, launches, dominant, split = [], [], [], 0, "k_submerged", []
        for _ in range(min(args.steps, 200)):
            step_device()
            t = ctx.timings()
            k_ms.append(max(t.submerged_ms, t.surface_ms))
            pro_ms.append(t.prologue_ms); epi_ms.append(t.epilogue_ms)
            launches = t.launches
            dominant = "k_submerged" if t.submerged_ms >= t.surface_ms else "surface path (k_classify + k_surface_lists: wet and clip roles)"
            split.append((t.classify_ms, t.wet_ms, t.clip_ms))
            if t.chunks_surface and not (t.chunks_wet or t.chunks_clip):  # direct mode: every warp classifies its own item, no lists
                chunks = {"surface": int(t.chunks_surface), "mode": "direct (no lists: kinds not recorded)"}
            else:
                chunks = {"surface": int(t.chunks_surface), "wet": int(t.chunks_wet), "clip": int(t.chunks_clip),
                          "dry": int(t.chunks_surface - t.chunks_wet - t.chunks_clip)}
        ctx.set_profiling(False)
        rig.barrier()

        # ---- e2e: host buffers through sfh_step, H2D + D2H inside the timed region (wall clock between barriers)
        for _ in range(args.warmup):
            step_host()
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        rig.barrier()
        e2e_s = time.perf_counter() - t0

    dev_ms, e2e_ms, kern_ms = rig.max_over_ranks([dev_ms, e2e_s * 1e3, float(np.mean(k_ms))])
    infos = h_infos.copy()  # results of the last e2e step: the DEBUG=0 kernels on the full-size workload
    res = {"workload": workload, "n": n, "first": first, "faces": faces, "verts": verts, "algo_bytes": algo_bytes, "dev_ms": dev_ms,
           "e2e_ms": e2e_ms, "kern_ms": kern_ms, "pro_ms": float(np.mean(pro_ms)), "epi_ms": float(np.mean(epi_ms)),
           "launches": int(launches), "dominant": dominant, "setup_s": t_setup,
           "surface_split": None if dominant == "k_submerged" else dict(zip(("classify_ms", "wet_ms", "clip_ms"), (float(x) for x in np.mean(split, axis=0))), chunks=chunks), "infos": infos, "scene": scene,
           "host_meshes": host_meshes, "h2d": int(n * BODY_STATE_DTYPE.itemsize), "d2h": int(n * INFO_DTYPE.itemsize + 4)}
    ctx.close()
    del d_states, d_infos
    torch.cuda.empty_cache()
    return res


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_obj(m: dict, peak: float, peak_src: str) -> dict:
    achieved = m["algo_bytes"] / (m["kern_ms"] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of the same kernel on the same
    # meshes (profiles/ncu_traffic_r0N.json: dram__bytes_read.sum + dram__bytes_write.sum), scaled by face count; null for
    # workloads that capture does not describe.
    traffic, traffic_src, secondary = None, None, None
    for name in ("ncu_traffic_r02.json", "ncu_traffic_r01.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if m["workload"] == "cfg4" and m["dominant"] == "k_submerged" and os.path.exists(tpath):
            t = json.load(open(tpath))
            traffic = int(t["dram_bytes_per_face"] * m["faces"])
            traffic_src = f"{t['dram_bytes_per_face']:.2f} B/face measured by ncu ({t['report']}) x {m['faces']} faces"
            secondary = dict(t.get("secondary") or {}, source=t["report"])  # the kernel is fp32-issue-bound, not HBM-bound
            break
    return {"bound": "hbm", "kernel": m["dominant"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(m["algo_bytes"]),
            "kernel_ms": m["kern_ms"], "prologue_ms": m["pro_ms"], "epilogue_ms": m["epi_ms"],
            "bytes_per_face": ALGO_BYTES_PER_FACE_NOTE, "secondary": secondary, "surface_split": m["surface_split"]}


def workload_record(m: dict, world: int, steps: int, peak: float, peak_src: str, validation: dict | None) -> dict:
    total_faces = m["faces"] * world
    return {"config": config_dict(m["workload"], m["n"]),
            "value": total_faces * steps / (m["dev_ms"] * 1e-3), "unit": UNIT, "ms_per_step": m["dev_ms"] / steps,
            "e2e": {"value": total_faces * steps / (m["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": m["e2e_ms"] / steps,
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "roofline": roofline_obj(m, peak, peak_src), "gpu_launches": m["launches"] * steps, "validation": validation,
            "setup_s": round(m["setup_s"], 2)}


def run_ours(args):
    from stonefish_b200.sharding import checksum, gather_infos, shard_range
    from tests.parity import COUNT_FIELDS, max_rel_err

    rig = Rig(args)
    rank, world, dev = rig.rank, rig.world, rig.dev
    _, per_gpu, _, _ = workload_spec(args.workload)
    n = args.bodies if args.bodies else per_gpu
    lead = rank == 0
    want_cpu = lead and world == 1 and not args.no_cpu
    keep = min(n, max(VALIDATE_BODIES, 4 * (os.cpu_count() or 1))) if (lead and not args.no_validate) else 0
    peak, peak_src = hbm_peak()

    sampler = ClockSampler(rig.local_rank)
    sampler.start()  # nvidia-smi needs ~0.1-0.3 s to start streaming: launch it before the upload and the warm-up
    head = measure(rig, args, args.workload, n, rank * n, keep, sampler)
    clocks = sampler.stop(trace=args.clock_trace)

    # ---- validation gather (C1), outside every timed region: per-body results + checksum over NCCL
    if world > 1:
        all_infos, sums = gather_infos(head["infos"], device=dev)
    else:
        all_infos, sums = head["infos"], [checksum(head["infos"]) & 0x7FFFFFFFFFFFFFFF]
    finite = bool(np.isfinite(all_infos["Fdq"]).all() and (all_infos["n_faces"] > 0).all())
    failures = [] if finite else ["non-finite or empty results"]

    line = None
    if lead:
        validation = {"checksums": [f"{s:016x}" for s in sums], "bodies_gathered": int(all_infos.shape[0]), "finite": finite}
        if keep:
            from oracle import pyoracle as po

            po.build()
            validation.update(validate_against_oracle(head["scene"], head["host_meshes"], head["infos"]))
            if not validation["ok"]:
                failures.append(f"{args.workload}: oracle mismatch {validation}")
        rec = workload_record(head, world, args.steps, peak, peak_src, None)
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 per-face terms, f64 accumulation", "data": "synthetic",
            "config": rec["config"], "setup_s": rec["setup_s"],
            # h2d = the body states (the ocean record, <= 6.2 KB, is only re-sent when it changes); d2h = the result records + status word
            "e2e": rec["e2e"],
            "gpu_launches": rec["gpu_launches"],
            "roofline": rec["roofline"],
            "clocks": clocks,
            "validation": validation,
        }

    # ---- strong scaling: BASELINE configs[3] as written — per_gpu instances IN TOTAL, block-partitioned over the ranks
    if world > 1 and args.workload in ("cfg4", "cfg5") and not args.bodies:
        lo, hi = shard_range(per_gpu, rank, world)
        sm = measure(rig, args, args.workload, hi - lo, lo, 0)
        s_all, _ = gather_infos(sm["infos"], device=dev)
        if lead:
            # rank 0's weak share is instances 0..per_gpu-1, i.e. exactly the strong job's bodies: same inputs, different partition
            err, where = max_rel_err(s_all, head["infos"])
            same_counts = all(np.array_equal(s_all[f], head["infos"][f]) for f in COUNT_FIELDS)
            ok = bool(err <= 1e-12 and same_counts)
            if not ok:
                failures.append(f"strong-scaling results differ from rank 0's: {err} {where} counts {same_counts}")
            total = sm["faces"]  # this rank's; shards are equal for 4096 / 128 over 2, 4, 8
            total_faces = per_gpu * (head["faces"] // n)
            line["strong"] = {
                "scaling": "strong", "bodies_total": per_gpu, "faces_total_per_step": total_faces,
                "shard_ranges": [list(shard_range(per_gpu, r, world)) for r in range(world)],
                "value": total_faces * args.steps / (sm["dev_ms"] * 1e-3), "unit": UNIT, "ms_per_step": sm["dev_ms"] / args.steps,
                "e2e": {"value": total_faces * args.steps / (sm["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": sm["e2e_ms"] / args.steps,
                        "h2d_bytes_per_step": sm["h2d"], "d2h_bytes_per_step": sm["d2h"]},
                "kernel_ms": sm["kern_ms"], "prologue_ms": sm["pro_ms"], "epilogue_ms": sm["epi_ms"],
                "fixed_cost_note": "ms_per_step - kernel_ms = prologue + epilogue + launch gaps: the part that does not shrink with N",
                "vs_rank0_full_run": {"max_rel_diff": err, "counts_equal": same_counts, "ok": ok},
            }
            del total

    # ---- the other BASELINE configs, same process, same rules
    extras = () if (args.no_extra or args.workload_given) else (EXTRA_WORKLOADS_1GPU if world == 1 else EXTRA_WORKLOADS_NGPU)
    if extras and lead:
        line["workloads"] = {}
    for w in extras:
        _, w_n, _, _ = workload_spec(w)
        keep_w = min(w_n, VALIDATE_BODIES) if (lead and not args.no_validate) else 0
        m = measure(rig, args, w, w_n, rank * w_n, keep_w)
        if lead:
            v = validate_against_oracle(m["scene"], m["host_meshes"], m["infos"]) if keep_w else None
            if v is not None and not v["ok"]:
                failures.append(f"{w}: oracle mismatch {v}")
            line["workloads"][w] = workload_record(m, world, args.steps, peak, peak_src, v)
            if want_cpu and args.cpu_extras:
                threads = os.cpu_count() or 1
                use = threads if w_n > 1 else 1
                hm = m["host_meshes"][:max(1, min(len(m["host_meshes"]), 2 * use))]
                rate, sample, _, _, kind = cpu_oracle_rate(w, m["scene"], hm, min(args.cpu_seconds, 4.0), use if len(hm) > 1 else 1)
                line["workloads"][w]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": use if len(hm) > 1 else 1, "kind": kind, "sample": sample}
        del m

    if lead:
        if want_cpu:
            from oracle import pyoracle as po

            po.build()
            threads = os.cpu_count() or 1
            hm = head["host_meshes"][:min(len(head["host_meshes"]), 4 * threads)]
            use = threads if len(hm) > 1 else 1
            rate, sample, _, _, kind = cpu_oracle_rate(args.workload, head["scene"], hm, args.cpu_seconds, use)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": use, "kind": kind, "sample": sample}
        line["ok"] = not failures
        if failures:
            line["failures"] = failures
        print(json.dumps(line), flush=True)
    code = rig.max_over_ranks([1.0 if failures else 0.0])[0]
    if world > 1:
        rig.dist.barrier(device_ids=[rig.local_rank])
        rig.dist.destroy_process_group()
    if failures:
        print("bench.py: VALIDATION FAILED: " + "; ".join(failures), file=sys.stderr, flush=True)
    return 1 if code else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=None, choices=["cfg4", "cfg4-shared", "cfg2", "cfg3", "cfg3-flat", "cfg5"],
                    help="headline workload (default cfg4, followed by the other BASELINE configs as `workloads`)")
    ap.add_argument("--bodies", type=int, default=0, help="bodies per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU time budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-extras", action="store_true", help="also time the CPU arm on each of the extra workloads")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    ap.add_argument("--no-validate", action="store_true", help="skip the oracle check of the timed results")
    ap.add_argument("--clock-trace", action="store_true", help="keep every nvidia-smi sample of the timed region in clocks.trace")
    args = ap.parse_args()
    args.workload_given = args.workload is not None
    if args.workload is None:
        args.workload = "cfg4"
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
