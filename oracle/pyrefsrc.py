"""ctypes binding of oracle/_ref/librefsrc.so — the reference's OWN hot-path definitions compiled here against a stand-in for
GLM (oracle/ref_harness.cpp, oracle/ref_shim/shim.hpp). TEST INFRASTRUCTURE ONLY; exists only where the reference checkout is
mounted (make -C oracle refsrc). Used to write tests/golden/ref_golden.json and by tests/test_oracle_refsource.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle.pyoracle import HERE, mesh_struct, orc_mesh
from stonefish_b200 import _abi
from stonefish_b200.hydro import INFO_DTYPE, Jet, Pipe, Uniform

LIB = os.path.join(HERE, "_ref", "librefsrc.so")


class refx_field(C.Structure):
    _fields_ = [("type", C.c_uint32), ("enabled", C.c_uint32), ("a", C.c_double * 3), ("b", C.c_double * 3), ("r1", C.c_double),
                ("r2", C.c_double), ("vel", C.c_double), ("gamma", C.c_double)]


class refx_fluid(C.Structure):
    _fields_ = [("density", C.c_double), ("gravity_z", C.c_double), ("has_waves", C.c_uint32), ("currents_enabled", C.c_uint32),
                ("n_fields", C.c_uint32), ("reserved", C.c_uint32), ("fields", refx_field * _abi.SFH_MAX_CURRENTS),
                ("fft_data", C.POINTER(C.c_float)), ("fft_size", C.c_int32), ("grid_sizes", C.c_float * 4)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        P = C.POINTER
        L.refx_depth_f.restype = C.c_float
        L.refx_depth_f.argtypes = [P(refx_fluid), P(C.c_float)]
        for n in ("refx_velocity_f", "refx_wind_f"):
            getattr(L, n).restype = None
            getattr(L, n).argtypes = [P(refx_fluid), P(C.c_float), P(C.c_float)]
        L.refx_velocity_d.restype = None
        L.refx_velocity_d.argtypes = [P(refx_fluid), P(C.c_double), P(C.c_double)]
        L.refx_body.restype = None
        L.refx_body.argtypes = [P(orc_mesh), P(refx_fluid), C.c_void_p, C.c_void_p]
        L.refx_step.restype = C.c_double
        L.refx_step.argtypes = [P(P(orc_mesh)), P(refx_fluid), C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.refx_aero.restype = None
        L.refx_aero.argtypes = [P(orc_mesh), P(refx_fluid), C.c_void_p, C.c_int, C.c_int, P(C.c_double), P(C.c_double), P(C.c_double)]
        L.refx_compound.restype = None
        L.refx_compound.argtypes = [P(P(orc_mesh)), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, P(refx_fluid), C.c_void_p,
                                    P(C.c_double), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refx_cable.restype = None
        L.refx_cable.argtypes = [P(refx_fluid), C.c_uint32, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refx_wave_tables.restype = None
        L.refx_wave_tables.argtypes = [P(_abi.sfh_wave_params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.refx_gl_matrix.restype = None
        L.refx_gl_matrix.argtypes = [P(C.c_double), P(C.c_float)]
        _lib = L
    return _lib


def fluid_struct(fields, density, gravity=0.0, currents_enabled=True, fft=None, fft_size=256, grid_sizes=(893.0, 101.0, 21.0, 11.0)):
    """refx_fluid from host-mirror field objects (their CONSTRUCTOR arguments: Jet.direction, Pipe.p2 are kept by hydro.py)."""
    f = refx_fluid()
    f.density, f.gravity_z = float(density), float(gravity)
    f.currents_enabled = 1 if currents_enabled else 0
    f.n_fields = len(fields)
    for i, c in enumerate(fields):
        d = f.fields[i]
        d.enabled = 1 if c.isEnabled() else 0
        if isinstance(c, Uniform):
            d.type, d.a[:] = 0, list(c.v)
        elif isinstance(c, Jet):
            d.type, d.a[:], d.b[:], d.r1, d.vel = 1, list(c.c), list(c.direction), c.r, c.vout
        else:
            assert isinstance(c, Pipe)
            d.type, d.a[:], d.b[:], d.r1, d.r2, d.vel, d.gamma = 2, list(c.p1), list(c.p2), c.r1, c.r2, c.vin, c.gamma
    if fft is not None:
        f.has_waves = 1
        f._keep = np.ascontiguousarray(fft, dtype=np.float32).reshape(-1)
        f.fft_data = f._keep.ctypes.data_as(C.POINTER(C.c_float))
        f.fft_size = fft_size
        f.grid_sizes[:] = list(grid_sizes)
    return f


def ocean_fluid(ocean):
    return fluid_struct(ocean.currents, ocean.density, ocean.gravity, ocean.currentsEnabled, ocean.fft_data if ocean.hasWaves() else None,
                        ocean.fft_size, ocean.grid_sizes)


def body(mesh, fluid, state: np.ndarray) -> np.ndarray:
    out = np.zeros(1, dtype=INFO_DTYPE)
    ms = mesh_struct(mesh)
    lib().refx_body(C.byref(ms), C.byref(fluid), np.ascontiguousarray(state).ctypes.data, out.ctypes.data)
    return out[0]


def aero(mesh, fluid, state: np.ndarray, dispatch: bool, approx_type: int, params):
    """(Fda, Tda) of SolidEntity::ComputeAerodynamicForces — static integrator, or the member incl. CorrectAerodynamicForces."""
    ms = mesh_struct(mesh)
    F, T = (C.c_double * 3)(), (C.c_double * 3)()
    lib().refx_aero(C.byref(ms), C.byref(fluid), np.ascontiguousarray(state).ctypes.data, 1 if dispatch else 0, int(approx_type),
                    (C.c_double * 3)(*[float(x) for x in params]), F, T)
    return np.array(F), np.array(T)


def compound(meshes, part_mesh_idx, parts: np.ndarray, external, origins: np.ndarray, o2c: np.ndarray, fluid, comp: np.ndarray, T_O: np.ndarray,
             aero_types=None, aero_params=None) -> np.ndarray:
    """Compound::ComputeHydrodynamicForces (or, with aero_types, ComputeAerodynamicForces) on the reference's own text."""
    ms = [mesh_struct(meshes[int(i)]) for i in part_mesh_idx]
    arr = (C.POINTER(orc_mesh) * len(ms))(*[C.pointer(m) for m in ms])
    ext = np.ascontiguousarray(np.asarray(external, dtype=np.uint32))
    parts = np.ascontiguousarray(parts)
    org = np.ascontiguousarray(origins, dtype=np.float64)
    oc = np.ascontiguousarray(o2c, dtype=np.float64)
    cs = np.ascontiguousarray(np.atleast_1d(comp))
    out = np.zeros(1, dtype=INFO_DTYPE)
    is_aero = aero_types is not None
    at = np.ascontiguousarray(np.asarray(aero_types if is_aero else [0] * len(ms), dtype=np.int32))
    ap = np.ascontiguousarray(np.asarray(aero_params if is_aero else np.ones((len(ms), 3)), dtype=np.float64))
    lib().refx_compound(arr, parts.ctypes.data, ext.ctypes.data, org.ctypes.data, oc.ctypes.data, len(ms), C.byref(fluid), cs.ctypes.data,
                        (C.c_double * 12)(*[float(x) for x in T_O]), 1 if is_aero else 0, at.ctypes.data, ap.ctypes.data, out.ctypes.data)
    return out[0]


def cable(fluid, radius: float, flags: int, pos: np.ndarray, vel: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((pos.shape[0], 9), dtype=np.float64)
    lib().refx_cable(C.byref(fluid), pos.shape[0], float(radius), int(flags), pos.ctypes.data, vel.ctypes.data, out.ctypes.data)
    return out


def depth(fluid, p) -> float:
    return float(lib().refx_depth_f(C.byref(fluid), (C.c_float * 3)(*[float(x) for x in p])))


def velocity_f(fluid, p) -> np.ndarray:
    out = (C.c_float * 3)()
    lib().refx_velocity_f(C.byref(fluid), (C.c_float * 3)(*[float(x) for x in p]), out)
    return np.array(out, dtype=np.float32)


def wind_f(fluid, p) -> np.ndarray:
    out = (C.c_float * 3)()
    lib().refx_wind_f(C.byref(fluid), (C.c_float * 3)(*[float(x) for x in p]), out)
    return np.array(out, dtype=np.float32)


def wave_tables(params, spectrum: bool = True):
    """(spectrum12, spectrum34, butterfly) of OpenGLOcean::GenerateWavesSpectrum / ComputeButterflyLookupTable. The reference's seed
    is a function-local static: ask for the spectrum only ONCE per process."""
    n = 1 << params.passes
    s12, s34 = np.zeros(n * n * 4, np.float32), np.zeros(n * n * 4, np.float32)
    bf = np.zeros(n * params.passes * 4, np.float32)
    lib().refx_wave_tables(C.byref(params), s12.ctypes.data if spectrum else None, s34.ctypes.data if spectrum else None, bf.ctypes.data)
    return s12, s34, bf


class RefScene:
    """A whole scene on the reference's own compiled code, stepped like the reference schedules it (one task per body on a pool,
    refx_step). Same role and interface as pyoracle.OracleScene.step: the CPU arm of bench.py."""

    def __init__(self, meshes, body_mesh, ocean, states: np.ndarray):
        self.meshes = meshes
        self.mstructs = [mesh_struct(m) for m in meshes]
        self.fluid = ocean_fluid(ocean)
        self.states = np.ascontiguousarray(states)
        arr = (C.POINTER(orc_mesh) * len(body_mesh))()
        for i, mi in enumerate(body_mesh):
            arr[i] = C.pointer(self.mstructs[int(mi)])
        self.mesh_ptrs = arr

    def step(self, n_threads: int = 1):
        out = np.zeros(self.states.shape[0], dtype=INFO_DTYPE)
        secs = lib().refx_step(self.mesh_ptrs, C.byref(self.fluid), self.states.ctypes.data, self.states.shape[0], n_threads, out.ctypes.data)
        return out, secs
