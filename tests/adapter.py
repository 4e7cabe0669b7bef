"""ctypes binding of oracle/_ref/libadapter.so (oracle/adapter_harness.cpp): the call-site adapter of
integration/stonefish_v1.6.0.patch next to the reference's own definitions. TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from oracle.pyoracle import mesh_struct, orc_mesh
from oracle.pyrefsrc import ocean_fluid, refx_fluid
from stonefish_b200.hydro import INFO_DTYPE

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libadapter.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        P = C.POINTER
        L.adp_new.restype = C.c_void_p
        L.adp_new.argtypes = [P(refx_fluid)]
        L.adp_free.argtypes = [C.c_void_p]
        L.adp_add_solid.restype = C.c_int
        L.adp_add_solid.argtypes = [C.c_void_p, P(orc_mesh), C.c_void_p]
        L.adp_add_compound.restype = C.c_int
        L.adp_add_compound.argtypes = [C.c_void_p, P(P(orc_mesh)), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, P(C.c_double)]
        L.adp_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, P(C.c_double)]
        L.adp_step.restype = C.c_int
        L.adp_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class AdapterScene:
    """A scene of sf::SolidEntity / sf::Compound objects stepped through the patched call site and through the reference text."""

    def __init__(self, ocean):
        self.fluid = ocean_fluid(ocean)
        self.h = lib().adp_new(C.byref(self.fluid))
        assert self.h, "adp_new failed (no CUDA device?)"
        self.keep, self.n = [], 0

    def add_solid(self, mesh, state):
        ms = mesh_struct(mesh)
        self.keep.append((mesh, ms))
        self.n += 1
        return lib().adp_add_solid(self.h, C.byref(ms), np.ascontiguousarray(state).ctypes.data)

    def add_compound(self, meshes, parts, external, origins, o2c, comp, T_O):
        ms = [mesh_struct(m) for m in meshes]
        arr = (C.POINTER(orc_mesh) * len(ms))(*[C.pointer(m) for m in ms])
        ext = np.ascontiguousarray(np.asarray(external, dtype=np.uint32))
        org, oc = np.ascontiguousarray(origins, dtype=np.float64), np.ascontiguousarray(o2c, dtype=np.float64)
        self.keep.append((meshes, ms))
        self.n += 1
        return lib().adp_add_compound(self.h, arr, np.ascontiguousarray(parts).ctypes.data, ext.ctypes.data, org.ctypes.data, oc.ctypes.data,
                                      len(ms), np.ascontiguousarray(np.atleast_1d(comp)).ctypes.data, (C.c_double * 12)(*[float(x) for x in T_O]))

    def set_state(self, k, state, T_O=None):
        lib().adp_set_state(self.h, k, np.ascontiguousarray(state).ctypes.data, (C.c_double * 12)(*[float(x) for x in T_O]) if T_O is not None else None)

    def step(self):
        g, r = np.zeros(self.n, dtype=INFO_DTYPE), np.zeros(self.n, dtype=INFO_DTYPE)
        assert lib().adp_step(self.h, g.ctypes.data, r.ctypes.data) == 0
        return g, r

    def close(self):
        if self.h:
            lib().adp_free(self.h)
            self.h = None
