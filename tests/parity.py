"""Shared comparison helpers of the parity tests: CUDA path (through the C ABI) vs the CPU oracle."""
import numpy as np

VEC_FIELDS = ("Fb", "Tb", "Fdq", "Tdq", "Fdf", "Tdf")
COUNT_FIELDS = ("class_count", "n_degenerate", "n_quadratic", "n_skin", "n_faces", "position", "wrote_buoyancy", "wrote_damping")

# north_star: "per-body Fb/Tb/Fd/Td/Fs/Ts must match within 1e-9 relative to the force norm"
REL_TOL = 1e-9


def max_rel_err(gpu: np.ndarray, ref: np.ndarray):
    """Largest |gpu - ref| / scale over bodies and outputs. Forces are scaled by the body's force norm
    (|Fb|+|Fdq|+|Fdf|), torques by the torque norm, Swet/Vsub by their own magnitude."""
    worst, where = 0.0, None
    F = sum(np.linalg.norm(ref[k], axis=-1) for k in ("Fb", "Fdq", "Fdf"))
    T = sum(np.linalg.norm(ref[k], axis=-1) for k in ("Tb", "Tdq", "Tdf"))
    for k in VEC_FIELDS:
        # north_star scales all six sums by the FORCE norm; a torque that vanishes by symmetry (a sphere floating at
        # x=y=0) is pure cancellation noise and must not be judged against its own ~0 magnitude.
        scale = F if k.startswith("F") else np.maximum(T, F)
        scale = np.maximum(scale, 1e-300)
        err = np.abs(gpu[k] - ref[k]).max(axis=-1) / scale
        err = np.where(np.abs(gpu[k] - ref[k]).max(axis=-1) == 0, 0.0, err)
        i = int(np.argmax(err))
        if err[i] > worst:
            worst, where = float(err[i]), (k, i)
    for k in ("Swet", "Vsub"):
        d = np.abs(gpu[k] - ref[k])
        err = np.where(d == 0, 0.0, d / np.maximum(np.abs(ref[k]), 1e-300))
        i = int(np.argmax(err))
        if err[i] > worst:
            worst, where = float(err[i]), (k, i)
    return worst, where


def assert_counts_equal(gpu: np.ndarray, ref: np.ndarray):
    for k in COUNT_FIELDS:
        assert np.array_equal(gpu[k], ref[k]), f"{k} differs: bodies {np.nonzero((gpu[k] != ref[k]).reshape(len(gpu), -1).any(axis=1))[0][:10]}"


def run_both(ctx, oracle, scene, debug=True, acc_mode=1):
    """(gpu_infos, gpu_codes, ref_infos, ref_codes) for one scene."""
    states = scene.bind(ctx)
    if debug:
        gi, gc = ctx.compute_debug(scene.ocean, states, scene.faces_per_step())
    else:
        gi, gc = ctx.compute(scene.ocean, states), None
    osc = oracle.OracleScene(scene.meshes, scene.body_mesh, scene.ocean, scene.states)
    if debug:
        ri, rc = osc.all_codes(acc_mode)
    else:
        ri, _ = osc.step(acc_mode, n_threads=8)
        rc = None
    return gi, gc, ri, rc
