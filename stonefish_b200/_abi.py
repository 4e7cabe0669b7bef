"""ctypes mirror of include/sfh_types.h and include/sfh.h, and the loader of libsfh.so.

The shared library is the product; this module only binds it. There is no Python/torch fallback: if
libsfh.so is missing or the CUDA device is unusable, loading or `sfh_create` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

SFH_MAX_CURRENTS = 64
SFH_CURRENT_UNIFORM, SFH_CURRENT_JET, SFH_CURRENT_PIPE = 0, 1, 2
SFH_POS_OUTSIDE, SFH_POS_INSIDE, SFH_POS_CROSSING, SFH_POS_AUTO = 0, 1, 2, 3
SFH_F_DAMPING, SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DISPATCH, SFH_F_AERO_CD = 1, 2, 4, 8, 16
SFH_MESH_KEEP_ORDER = 1
SFH_HINT_NO_SURFACE, SFH_HINT_NO_SUBMERGED = 1, 2
SFH_FACE_CLASS_MASK, SFH_FACE_DEGENERATE, SFH_FACE_QUADRATIC, SFH_FACE_SKIN = 7, 8, 16, 32
SFH_OK = 0


class sfh_current(C.Structure):
    _fields_ = [("type", C.c_uint32), ("enabled", C.c_uint32), ("a", C.c_double * 3), ("n", C.c_double * 3),
                ("r1", C.c_double), ("r2", C.c_double), ("l", C.c_double), ("vel", C.c_double), ("gamma", C.c_double)]


class sfh_ocean(C.Structure):
    _fields_ = [("density", C.c_double), ("gravity_z", C.c_double), ("has_waves", C.c_uint32),
                ("currents_enabled", C.c_uint32), ("n_currents", C.c_uint32), ("reserved", C.c_uint32),
                ("currents", sfh_current * SFH_MAX_CURRENTS)]


class sfh_atmosphere(C.Structure):
    _fields_ = [("density", C.c_double), ("n_winds", C.c_uint32), ("reserved", C.c_uint32),
                ("winds", sfh_current * SFH_MAX_CURRENTS)]


class sfh_waves(C.Structure):
    _fields_ = [("fft_size", C.c_int32), ("grid_sizes", C.c_float * 4), ("data", C.POINTER(C.c_float)),
                ("n_floats", C.c_size_t)]


class sfh_wave_params(C.Structure):
    _fields_ = [("wind", C.c_float), ("omega", C.c_float), ("A", C.c_float), ("km", C.c_float), ("cm", C.c_float),
                ("propagate", C.c_uint32), ("passes", C.c_uint32), ("grid_sizes", C.c_float * 4), ("reserved", C.c_uint32),
                ("seed", C.c_int64)]


class sfh_body_state(C.Structure):
    _fields_ = [("mesh_id", C.c_uint32), ("mode", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("T_CG", C.c_double * 12), ("T_C", C.c_double * 12), ("v", C.c_double * 3), ("omega", C.c_double * 3),
                ("aabb_min", C.c_double * 3), ("aabb_max", C.c_double * 3), ("T_O_basis", C.c_double * 9),
                ("Cd", C.c_double * 3), ("Cf", C.c_double * 3), ("volume", C.c_double), ("surface", C.c_double),
                ("P_CB", C.c_double * 3)]


class sfh_cable(C.Structure):
    _fields_ = [("first_node", C.c_uint32), ("n_nodes", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("radius", C.c_double)]


class sfh_node_forces(C.Structure):
    _fields_ = [("Fb", C.c_double * 3), ("Fdq", C.c_double * 3), ("Fdf", C.c_double * 3)]


class sfh_info(C.Structure):
    _fields_ = [("Fb", C.c_double * 3), ("Tb", C.c_double * 3), ("Fdq", C.c_double * 3), ("Tdq", C.c_double * 3),
                ("Fdf", C.c_double * 3), ("Tdf", C.c_double * 3), ("Swet", C.c_double), ("Vsub", C.c_double),
                ("wrote_buoyancy", C.c_uint32), ("wrote_damping", C.c_uint32), ("position", C.c_uint32),
                ("n_faces", C.c_uint32), ("class_count", C.c_uint32 * 8), ("n_degenerate", C.c_uint32),
                ("n_quadratic", C.c_uint32), ("n_skin", C.c_uint32), ("reserved", C.c_uint32)]


class sfh_mesh_stats(C.Structure):
    _fields_ = [("n_faces", C.c_uint32), ("n_verts", C.c_uint32), ("n_chunks", C.c_uint32), ("n_chunk_verts", C.c_uint32),
                ("device_bytes", C.c_uint64), ("algorithmic_bytes", C.c_uint64)]


class sfh_timings(C.Structure):
    _fields_ = [("prologue_ms", C.c_float), ("submerged_ms", C.c_float), ("surface_ms", C.c_float),
                ("epilogue_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32), ("reserved", C.c_uint32),
                ("faces", C.c_uint64), ("classify_ms", C.c_float), ("wet_ms", C.c_float), ("clip_ms", C.c_float),
                ("chunks_surface", C.c_uint32), ("chunks_wet", C.c_uint32), ("chunks_clip", C.c_uint32)]


# numpy structured dtypes with the same layout (used for vectorised scene building and comparisons)
def np_dtype(struct):
    import numpy as np

    fields = []
    for name, ctype in struct._fields_:
        if hasattr(ctype, "_length_") and not issubclass(ctype, C.Structure):
            base = np.dtype(ctype._type_)
            fields.append((name, base, (ctype._length_,)))
        else:
            fields.append((name, np.dtype(ctype)))
    dt = np.dtype(fields, align=True)
    assert dt.itemsize == C.sizeof(struct), (struct.__name__, dt.itemsize, C.sizeof(struct))
    return dt


LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsfh.so")

# every symbol include/sfh.h declares, with its prototype
_P = C.POINTER
_ctx = C.c_void_p
PROTOTYPES = {
    "sfh_abi_version": (C.c_int, []),
    "sfh_create": (C.c_int, [C.c_int, _P(_ctx)]),
    "sfh_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, _P(_ctx)]),
    "sfh_destroy": (None, [_ctx]),
    "sfh_last_error": (C.c_char_p, [_ctx]),
    "sfh_mesh_upload": (C.c_int, [_ctx, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, _P(C.c_uint32)]),
    "sfh_mesh_release": (C.c_int, [_ctx, C.c_uint32]),
    "sfh_mesh_get_stats": (C.c_int, [_ctx, C.c_uint32, _P(sfh_mesh_stats)]),
    "sfh_waves_upload": (C.c_int, [_ctx, _P(sfh_waves)]),
    "sfh_wavegen_init": (C.c_int, [_ctx, _P(sfh_wave_params)]),
    "sfh_wavegen_simulate": (C.c_int, [_ctx, C.c_float]),
    "sfh_wavegen_download": (C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    "sfh_wavegen_get_spectrum": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sfh_step_async": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_sync": (C.c_int, [_ctx]),
    "sfh_step": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_step_debug": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sfh_step_device": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_step_device_hinted": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]),
    "sfh_aero_step": (C.c_int, [_ctx, _P(sfh_atmosphere), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_aero_step_debug": (C.c_int, [_ctx, _P(sfh_atmosphere), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sfh_aero_step_device": (C.c_int, [_ctx, _P(sfh_atmosphere), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_cable_step": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_query_depth": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_query_fluid_velocity": (C.c_int, [_ctx, _P(sfh_ocean), C.c_void_p, C.c_uint32, C.c_void_p]),
    "sfh_selftest": (C.c_int, [_ctx, C.c_uint32, C.c_uint64, _P(C.c_uint64), _P(C.c_uint64)]),
    "sfh_host_alloc": (C.c_int, [_ctx, C.c_size_t, _P(C.c_void_p)]),
    "sfh_host_free": (C.c_int, [_ctx, C.c_void_p]),
    "sfh_set_profiling": (C.c_int, [_ctx, C.c_int]),
    "sfh_get_timings": (C.c_int, [_ctx, _P(sfh_timings)]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libsfh.so and attach prototypes. Raises if the library was not built (run
    `python -c "import __graft_entry__ as g; g.build()"` or `make -C stonefish_b200/csrc`)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SFH_LIB") or LIB_PATH  # SFH_LIB: A/B builds during kernel tuning
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: the CUDA extension is not built and there is no CPU fallback "
                           "(make -C stonefish_b200/csrc)")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.sfh_abi_version() != 2:
        raise RuntimeError("libsfh.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
