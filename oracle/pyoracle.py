"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package. Pinned against the reference's own source text except for the GLM primitives (see oracle/hydro_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from stonefish_b200 import _abi
from stonefish_b200.hydro import BODY_STATE_DTYPE, INFO_DTYPE, Mesh, Ocean

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libbtcheck.so")
ORC_REF_F32_SEQ, ORC_F64_ACC = 0, 1


class orc_mesh(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("stride_bytes", C.c_size_t), ("n_verts", C.c_uint32), ("faces", C.c_void_p),
                ("n_faces", C.c_uint32)]


def build(force: bool = False):
    srcs = [os.path.join(HERE, f) for f in ("hydro_oracle.cpp", "wave_oracle.cpp", "hydro_oracle.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/3rdparty/LinearMath"):
        subprocess.check_call(["make", "-C", HERE, "ref", "refsrc"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        P = C.POINTER
        D3 = P(C.c_double)
        L.orc_wave_bilinear.restype = C.c_float
        L.orc_wave_bilinear.argtypes = [P(_abi.sfh_waves), C.c_float, C.c_float, C.c_uint32]
        L.orc_wave_height.restype = C.c_float
        L.orc_wave_height.argtypes = [P(_abi.sfh_waves), C.c_float, C.c_float]
        L.orc_get_depth.restype = C.c_float
        L.orc_get_depth.argtypes = [P(_abi.sfh_ocean), P(_abi.sfh_waves), C.c_float, C.c_float, C.c_float]
        L.orc_current_velocity.restype = None
        L.orc_current_velocity.argtypes = [P(_abi.sfh_current), D3, D3]
        L.orc_fluid_velocity_d.restype = None
        L.orc_fluid_velocity_d.argtypes = [P(_abi.sfh_ocean), D3, D3]
        L.orc_fluid_velocity_f.restype = None
        L.orc_fluid_velocity_f.argtypes = [P(_abi.sfh_ocean), P(C.c_float), P(C.c_float)]
        L.orc_check_body_fluid_position.restype = C.c_uint32
        L.orc_check_body_fluid_position.argtypes = [P(_abi.sfh_ocean), P(_abi.sfh_waves), D3, D3]
        L.orc_correct.restype = None
        L.orc_correct.argtypes = [C.c_double, D3, D3, D3, D3, D3, D3, D3]
        L.orc_gl_matrix.restype = None
        L.orc_gl_matrix.argtypes = [D3, P(C.c_float)]
        for name in ("orc_submerged",):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [P(orc_mesh), P(_abi.sfh_ocean), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("orc_surface", "orc_body"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [P(orc_mesh), P(_abi.sfh_ocean), P(_abi.sfh_waves), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_step.restype = C.c_double
        L.orc_step.argtypes = [P(P(orc_mesh)), P(_abi.sfh_ocean), P(_abi.sfh_waves), C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.orc_body_split.restype = C.c_double
        L.orc_body_split.argtypes = [P(orc_mesh), P(_abi.sfh_ocean), P(_abi.sfh_waves), C.c_void_p, C.c_int, C.c_void_p]
        L.orc_selftest_unfused.restype = C.c_int
        _lib = L
    return _lib


def mesh_struct(m: Mesh) -> orc_mesh:
    return orc_mesh(m.vertex_data.ctypes.data, m.stride_bytes, m.n_verts, m.faces.ctypes.data, m.n_faces)


class OracleScene:
    """Keeps the ctypes views of a scene alive and runs oracle entry points on it."""

    def __init__(self, meshes: list[Mesh], body_mesh, ocean: Ocean, states: np.ndarray):
        self.meshes = meshes
        self.body_mesh = np.asarray(body_mesh)
        self.mstructs = [mesh_struct(m) for m in meshes]
        self.ocean = ocean
        self.o = ocean.to_struct()
        self.w = ocean.waves_struct() if ocean.hasWaves() else _abi.sfh_waves()
        self.states = np.ascontiguousarray(states)
        assert self.states.dtype == BODY_STATE_DTYPE
        arr = (C.POINTER(orc_mesh) * len(self.body_mesh))()
        for i, mi in enumerate(self.body_mesh):
            arr[i] = C.pointer(self.mstructs[int(mi)])
        self.mesh_ptrs = arr

    def body(self, i: int, acc_mode: int = ORC_F64_ACC, codes: bool = False):
        out = np.zeros(1, dtype=INFO_DTYPE)
        m = self.meshes[int(self.body_mesh[i])]
        cbuf = np.zeros(max(m.n_faces, 1), dtype=np.uint8) if codes else None
        lib().orc_body(C.byref(self.mstructs[int(self.body_mesh[i])]), C.byref(self.o), C.byref(self.w),
                       self.states[i:i + 1].ctypes.data, acc_mode, out.ctypes.data, cbuf.ctypes.data if codes else None)
        return (out[0], cbuf[: m.n_faces]) if codes else out[0]

    def step(self, acc_mode: int = ORC_F64_ACC, n_threads: int = 1, first: int = 0, count: int | None = None):
        n = len(self.body_mesh) - first if count is None else count
        out = np.zeros(n, dtype=INFO_DTYPE)
        ptrs = C.cast(C.byref(self.mesh_ptrs, first * C.sizeof(C.c_void_p)), C.POINTER(C.POINTER(orc_mesh)))
        secs = lib().orc_step(ptrs, C.byref(self.o), C.byref(self.w), self.states[first:first + n].ctypes.data, n, acc_mode,
                              n_threads, out.ctypes.data)
        return out, secs

    def body_split(self, i: int, n_threads: int):
        out = np.zeros(1, dtype=INFO_DTYPE)
        secs = lib().orc_body_split(C.byref(self.mstructs[int(self.body_mesh[i])]), C.byref(self.o), C.byref(self.w),
                                    self.states[i:i + 1].ctypes.data, n_threads, out.ctypes.data)
        return out[0], secs

    def all_codes(self, acc_mode: int = ORC_F64_ACC):
        infos, codes = [], []
        for i in range(len(self.body_mesh)):
            info, c = self.body(i, acc_mode, codes=True)
            infos.append(info)
            codes.append(c)
        return np.array(infos, dtype=INFO_DTYPE), np.concatenate(codes) if codes else np.zeros(0, np.uint8)


def compound(meshes, part_mesh_idx, parts: np.ndarray, external, ocean: Ocean, compound_state: np.ndarray, acc_mode: int = ORC_F64_ACC):
    """Compound::ComputeHydrodynamicForces through the oracle (orc_compound)."""
    L = lib()
    if not hasattr(L, "_compound_ready"):
        L.orc_compound.restype = None
        L.orc_compound.argtypes = [C.POINTER(C.POINTER(orc_mesh)), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(_abi.sfh_ocean),
                                   C.POINTER(_abi.sfh_waves), C.c_void_p, C.c_int, C.c_void_p]
        L._compound_ready = True
    ms = [mesh_struct(meshes[int(i)]) for i in part_mesh_idx]
    arr = (C.POINTER(orc_mesh) * len(ms))(*[C.pointer(m) for m in ms])
    o = ocean.to_struct()
    w = ocean.waves_struct() if ocean.hasWaves() else _abi.sfh_waves()
    ext = np.ascontiguousarray(np.asarray(external, dtype=np.uint32))
    parts = np.ascontiguousarray(parts)
    cs = np.ascontiguousarray(np.atleast_1d(compound_state))
    out = np.zeros(1, dtype=INFO_DTYPE)
    L.orc_compound(arr, parts.ctypes.data, ext.ctypes.data, len(ms), C.byref(o), C.byref(w), cs.ctypes.data, acc_mode, out.ctypes.data)
    return out[0]


def _aero_lib():
    L = lib()
    if not hasattr(L, "_aero_ready"):
        P, D3 = C.POINTER, C.POINTER(C.c_double)
        L.orc_wind_velocity_f.restype = None
        L.orc_wind_velocity_f.argtypes = [P(_abi.sfh_atmosphere), P(C.c_float), P(C.c_float)]
        L.orc_correct_aero.restype = None
        L.orc_correct_aero.argtypes = [D3, D3, D3, D3, C.c_int]
        L.orc_aero.restype = None
        L.orc_aero.argtypes = [P(orc_mesh), P(_abi.sfh_atmosphere), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_compound_aero.restype = None
        L.orc_compound_aero.argtypes = [P(P(orc_mesh)), C.c_void_p, C.c_void_p, C.c_uint32, P(_abi.sfh_atmosphere), C.c_void_p,
                                        C.c_int, C.c_void_p]
        L._aero_ready = True
    return L


def aero(meshes, body_mesh, atm, states: np.ndarray, acc_mode: int = ORC_F64_ACC, codes: bool = False):
    """SolidEntity::ComputeAerodynamicForces for every body through the oracle (orc_aero); same contract as sfh_aero_step."""
    L = _aero_lib()
    a = atm.to_struct()
    states = np.ascontiguousarray(states)
    out = np.zeros(states.shape[0], dtype=INFO_DTYPE)
    all_codes = []
    for i in range(states.shape[0]):
        m = meshes[int(body_mesh[i])]
        ms = mesh_struct(m)
        cbuf = np.zeros(max(m.n_faces, 1), dtype=np.uint8) if codes else None
        L.orc_aero(C.byref(ms), C.byref(a), states[i:i + 1].ctypes.data, acc_mode, out[i:i + 1].ctypes.data,
                   cbuf.ctypes.data if codes else None)
        if codes:
            all_codes.append(cbuf[: m.n_faces])
    if codes:
        return out, (np.concatenate(all_codes) if all_codes else np.zeros(0, np.uint8))
    return out


def compound_aero(meshes, part_mesh_idx, parts: np.ndarray, external, atm, compound_state: np.ndarray, acc_mode: int = ORC_F64_ACC):
    """Compound::ComputeAerodynamicForces through the oracle (orc_compound_aero)."""
    L = _aero_lib()
    ms = [mesh_struct(meshes[int(i)]) for i in part_mesh_idx]
    arr = (C.POINTER(orc_mesh) * len(ms))(*[C.pointer(m) for m in ms])
    a = atm.to_struct()
    ext = np.ascontiguousarray(np.asarray(external, dtype=np.uint32))
    parts = np.ascontiguousarray(parts)
    cs = np.ascontiguousarray(np.atleast_1d(compound_state))
    out = np.zeros(1, dtype=INFO_DTYPE)
    L.orc_compound_aero(arr, parts.ctypes.data, ext.ctypes.data, len(ms), C.byref(a), cs.ctypes.data, acc_mode, out.ctypes.data)
    return out[0]


class WaveOracle:
    """OpenGLOcean's wave-field producer through the oracle (oracle/wave_oracle.cpp): spectrum + butterfly table once,
    then `simulate(t)` = oceanInit.frag + 16 FFT passes -> fftData [4, N, N, 4]."""

    def __init__(self, params: "_abi.sfh_wave_params"):
        L = lib()
        if not hasattr(L, "_wave_ready"):
            P = C.POINTER
            L.orc_half_round.restype = C.c_float
            L.orc_half_round.argtypes = [C.c_float]
            L.orc_wave_spectrum.restype = None
            L.orc_wave_spectrum.argtypes = [P(_abi.sfh_wave_params), C.c_void_p, C.c_void_p]
            L.orc_wave_butterfly.restype = None
            L.orc_wave_butterfly.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
            L.orc_wave_simulate.restype = None
            L.orc_wave_simulate.argtypes = [P(_abi.sfh_wave_params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
            L._wave_ready = True
        self.L = L
        self.params = params
        self.N = 1 << params.passes
        self.spectrum12 = np.zeros(self.N * self.N * 4, dtype=np.float32)
        self.spectrum34 = np.zeros(self.N * self.N * 4, dtype=np.float32)
        L.orc_wave_spectrum(C.byref(params), self.spectrum12.ctypes.data, self.spectrum34.ctypes.data)
        self.butterfly = np.zeros(self.N * params.passes * 4, dtype=np.float32)
        L.orc_wave_butterfly(self.N, params.passes, self.butterfly.ctypes.data)

    def simulate(self, t: float, half_storage: bool = True, spectrum12=None, spectrum34=None) -> np.ndarray:
        s12 = self.spectrum12 if spectrum12 is None else np.ascontiguousarray(spectrum12, dtype=np.float32)
        s34 = self.spectrum34 if spectrum34 is None else np.ascontiguousarray(spectrum34, dtype=np.float32)
        out = np.zeros((4, self.N, self.N, 4), dtype=np.float32)
        self.L.orc_wave_simulate(C.byref(self.params), s12.ctypes.data, s34.ctypes.data, self.butterfly.ctypes.data, float(t),
                                 1 if half_storage else 0, out.ctypes.data)
        return out


def cable(ocean: Ocean, radius: float, flags: int, pos: np.ndarray, vel: np.ndarray) -> np.ndarray:
    """CableEntity::ComputeHydrodynamicForces for one cable through the oracle (orc_cable): (n_nodes, 9) = Fb, Fdq, Fdf."""
    L = lib()
    if not hasattr(L, "_cable_ready"):
        L.orc_cable.restype = None
        L.orc_cable.argtypes = [C.POINTER(_abi.sfh_ocean), C.POINTER(_abi.sfh_waves), C.POINTER(_abi.sfh_cable), C.c_void_p, C.c_void_p,
                                C.c_void_p]
        L.orc_circular_segment_area.restype = C.c_double
        L.orc_circular_segment_area.argtypes = [C.c_double, C.c_double]
        L._cable_ready = True
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(-1, 3)
    c = _abi.sfh_cable(0, pos.shape[0], flags, 0, float(radius))
    o = ocean.to_struct()
    w = ocean.waves_struct() if ocean.hasWaves() else _abi.sfh_waves()
    out = np.zeros((pos.shape[0], 9), dtype=np.float64)
    L.orc_cable(C.byref(o), C.byref(w), C.byref(c), pos.ctypes.data, vel.ctypes.data, out.ctypes.data)
    return out
