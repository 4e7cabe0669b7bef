"""Host-side mirror of the reference interface for the hydrodynamics path, over the C ABI (libsfh.so).

Names follow the reference (Library/include/entities/forcefields/{Ocean,Uniform,Jet,Pipe}.h,
Library/include/entities/SolidEntity.h): `Ocean.AddVelocityField`, `EnableCurrents`, `HydrodynamicsSettings`
with its (misspelt) `reallisticBuoyancy`, `BodyFluidPosition`. The C++ twin of this module — the one a Stonefish
build links — is stonefish_b200/host/StonefishHydro.hpp; this Python one drives tests and bench.py.

All compute happens in the CUDA library. Nothing here evaluates a face.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import (SFH_CURRENT_JET, SFH_CURRENT_PIPE, SFH_CURRENT_UNIFORM, SFH_F_AERO_CD, SFH_F_BUOYANCY, SFH_F_BUOYANT,
                   SFH_F_DAMPING, SFH_F_DISPATCH, SFH_POS_AUTO, SFH_POS_CROSSING, SFH_POS_INSIDE, SFH_POS_OUTSIDE)

BODY_STATE_DTYPE = _abi.np_dtype(_abi.sfh_body_state)
INFO_DTYPE = _abi.np_dtype(_abi.sfh_info)


class BodyFluidPosition:  # sf::BodyFluidPosition, SolidEntity.h:41
    OUTSIDE, INSIDE, CROSSING_SURFACE, AUTO = SFH_POS_OUTSIDE, SFH_POS_INSIDE, SFH_POS_CROSSING, SFH_POS_AUTO


@dataclass
class HydrodynamicsSettings:  # Ocean.h:36-40
    dampingForces: bool = True
    reallisticBuoyancy: bool = True

    def flags(self) -> int:
        return (SFH_F_DAMPING if self.dampingForces else 0) | (SFH_F_BUOYANCY if self.reallisticBuoyancy else 0)


# ------------------------------------------------------------------------------------------------ currents
class VelocityField:
    def __init__(self):
        self.enabled = True  # VelocityField.cpp:31

    def setEnabled(self, en: bool):
        self.enabled = bool(en)

    def isEnabled(self) -> bool:
        return self.enabled

    def _fill(self, c: _abi.sfh_current):
        raise NotImplementedError


def _v3(x):
    a = np.asarray(x, dtype=np.float64).reshape(3)
    return a


class Uniform(VelocityField):  # Uniform.cpp:31-57
    def __init__(self, velocity):
        super().__init__()
        self.v = _v3(velocity)

    def setVelocity(self, x):
        self.v = _v3(x)

    def _fill(self, c):
        c.type = SFH_CURRENT_UNIFORM
        c.a[:] = self.v.tolist()


class Jet(VelocityField):  # Jet.cpp:31-75
    def __init__(self, point, direction, radius, outletVelocity):
        super().__init__()
        d = _v3(direction)
        self.direction = d.copy()  # constructor argument, kept so scenes can be replayed through other bindings
        self.c = _v3(point)
        # btVector3::normalized(): *this / length() == *this * (1.0 / sqrt(x*x + y*y + z*z))
        self.n = d * (1.0 / np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        self.r = float(radius)
        self.vout = float(outletVelocity)

    def setOutletVelocity(self, x):
        self.vout = float(x)

    def _fill(self, c):
        c.type = SFH_CURRENT_JET
        c.a[:] = self.c.tolist()
        c.n[:] = self.n.tolist()
        c.r1 = self.r
        c.vel = self.vout


class Pipe(VelocityField):  # Pipe.cpp:31-78
    def __init__(self, point1, point2, radius1, radius2, inletVelocity, exponent):
        super().__init__()
        p1, p2 = _v3(point1), _v3(point2)
        self.p2 = p2.copy()
        d = p2 - p1
        self.p1 = p1
        self.l = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        self.n = d * (1.0 / self.l)
        self.r1, self.r2 = float(radius1), float(radius2)
        self.vin, self.gamma = float(inletVelocity), float(exponent)

    def setInletVelocity(self, x):
        self.vin = float(x)

    def _fill(self, c):
        c.type = SFH_CURRENT_PIPE
        c.a[:] = self.p1.tolist()
        c.n[:] = self.n.tolist()
        c.r1, c.r2, c.l, c.vel, c.gamma = self.r1, self.r2, self.l, self.vin, self.gamma


# ------------------------------------------------------------------------------------------------ ocean
class Ocean:
    """What the integrators read from sf::Ocean (Ocean.cpp:139-211): liquid density, gravity, the list of
    velocity fields, `currentsEnabled`, and the wave field (`fftData`, an input here)."""

    _wave_token = 0

    def __init__(self, density: float = 1000.0, gravity: float = 9.81):
        self.density = float(density)
        self.gravity = float(gravity)
        self.currents: list[VelocityField] = []
        self.currentsEnabled = False
        self.fft_data: np.ndarray | None = None
        self.fft_size = 256
        self.grid_sizes = (893.0, 101.0, 21.0, 11.0)  # OpenGLOcean.cpp:70
        self.waves_version = 0  # process-unique token of the current wave field (0 = none)
        self.generated_waves = False  # wave field produced on the device by HydroContext.wavegen_simulate

    def AddVelocityField(self, field: VelocityField):
        self.currents.append(field)

    def EnableCurrents(self):
        self.currentsEnabled = True

    def DisableCurrents(self):
        self.currentsEnabled = False

    def setWaves(self, fft_data: np.ndarray | None, fft_size: int = 256, grid_sizes=(893.0, 101.0, 21.0, 11.0)):
        """fft_data: OpenGLRealOcean::fftData, float32 [layers, N, N, 4] (at least layer 0 + N+1 texels)."""
        if fft_data is None:
            self.fft_data = None
        else:
            self.fft_data = np.ascontiguousarray(fft_data, dtype=np.float32).reshape(-1)
            self.fft_size = int(fft_size)
            self.grid_sizes = tuple(float(g) for g in grid_sizes)
        Ocean._wave_token += 1
        self.waves_version = Ocean._wave_token

    def useGeneratedWaves(self, on: bool = True):
        """The wave field is the one HydroContext.wavegen_simulate left on the device (no fftData upload)."""
        self.generated_waves = bool(on)

    def hasWaves(self) -> bool:
        return self.fft_data is not None or self.generated_waves

    def to_struct(self) -> _abi.sfh_ocean:
        o = _abi.sfh_ocean()
        o.density, o.gravity_z = self.density, self.gravity
        o.has_waves = 1 if self.hasWaves() else 0
        o.currents_enabled = 1 if self.currentsEnabled else 0
        if len(self.currents) > _abi.SFH_MAX_CURRENTS:
            raise ValueError("more than 64 velocity fields")
        o.n_currents = len(self.currents)
        for i, f in enumerate(self.currents):
            o.currents[i].enabled = 1 if f.isEnabled() else 0
            f._fill(o.currents[i])
        return o

    def waves_struct(self) -> _abi.sfh_waves:
        w = _abi.sfh_waves()
        w.fft_size = self.fft_size
        w.grid_sizes[:] = list(self.grid_sizes)
        w.data = self.fft_data.ctypes.data_as(C.POINTER(C.c_float))
        w.n_floats = self.fft_data.size
        return w


# ------------------------------------------------------------------------------------------------ meshes
class GeometryApproxType:  # sf::GeometryApproxType, SolidEntity.h:39
    AUTO, SPHERE, CYLINDER, ELLIPSOID = 0, 1, 2, 3


class Atmosphere:
    """What the aerodynamic integrator reads from sf::Atmosphere (Atmosphere.cpp:72-75,117-157): the gas density and the
    wind list, which the reference sums without consulting isEnabled()."""

    def __init__(self, density: float = 1.2):
        self.density = float(density)
        self.wind: list[VelocityField] = []

    def AddVelocityField(self, field: VelocityField):
        self.wind.append(field)

    def to_struct(self) -> _abi.sfh_atmosphere:
        a = _abi.sfh_atmosphere()
        a.density = self.density
        if len(self.wind) > _abi.SFH_MAX_CURRENTS:
            raise ValueError("more than 64 velocity fields")
        a.n_winds = len(self.wind)
        for i, f in enumerate(self.wind):
            a.winds[i].enabled = 1 if f.isEnabled() else 0
            f._fill(a.winds[i])
        return a


def aero_proxy(approx_type: int, params=()):
    """(Cd vector, flag) that CorrectAerodynamicForces derives from fdApproxType / fdApproxParams
    (SolidEntity.cpp:1931-1951), in the reference's own operation order (btVector3::operator/= multiplies by 1/s)."""
    if approx_type == GeometryApproxType.CYLINDER:
        return np.array([0.5, 0.5, 1.0]), SFH_F_AERO_CD
    if approx_type == GeometryApproxType.ELLIPSOID:
        cd = np.array([1.0 / float(params[0]), 1.0 / float(params[1]), 1.0 / float(params[2])])
        return cd * (1.0 / max(max(cd[0], cd[1]), cd[2])), SFH_F_AERO_CD
    return np.zeros(3), 0


def _libm_expf(x: float) -> float:
    """The C library's expf, so that parameters derived here equal the ones the C++ mirror derives (numpy's float32 exp is a
    different implementation and may differ in the last bit)."""
    import ctypes.util

    m = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    m.expf.restype, m.expf.argtypes = C.c_float, [C.c_float]
    return float(m.expf(x))


def wave_params(sea_state: float, seed: int = 1234) -> _abi.sfh_wave_params:
    """OceanParams as OpenGLOcean / OpenGLRealOcean set them for a sea state (OpenGLOcean.cpp:62-72,
    OpenGLRealOcean.cpp:41-45), in the reference's float arithmetic."""
    f = np.float32
    p = _abi.sfh_wave_params()
    p.wind = float(f(sea_state) * f(5.0) + f(2.0))
    p.omega = float(f(5.0) * f(_libm_expf(-float(f(sea_state)))) + f(0.2))
    p.A, p.km, p.cm = 1.0, 370.0, 0.23
    p.propagate, p.passes = 1, 8
    p.grid_sizes[:] = [893.0, 101.0, 21.0, 11.0]
    p.seed = seed
    return p


class Mesh:
    """sf::Mesh as the integrators see it (OpenGLDataStructs.h:120-262): vertex positions (possibly strided inside
    a Vertex / TexturableVertex array) and 3 vertex ids per face."""

    def __init__(self, vertices: np.ndarray, faces: np.ndarray, vertex_stride_floats: int = 3):
        v = np.asarray(vertices, dtype=np.float32)
        assert v.ndim == 2 and v.shape[1] >= 3
        if vertex_stride_floats != v.shape[1]:
            padded = np.zeros((v.shape[0], vertex_stride_floats), dtype=np.float32)
            padded[:, : v.shape[1]] = v
            v = padded
        self.vertex_data = np.ascontiguousarray(v)  # [V, stride] – pos in the first 3 floats, like Vertex{pos,normal}
        self.faces = np.ascontiguousarray(np.asarray(faces, dtype=np.uint32).reshape(-1, 3))

    @property
    def n_verts(self):
        return self.vertex_data.shape[0]

    @property
    def n_faces(self):
        return self.faces.shape[0]

    @property
    def stride_bytes(self):
        return self.vertex_data.shape[1] * 4

    @property
    def positions(self):
        return self.vertex_data[:, :3]


# ------------------------------------------------------------------------------------------------ transforms
def quat_to_basis(q) -> np.ndarray:
    """btMatrix3x3::setRotation (3rdparty/LinearMath/btMatrix3x3.h:211-225), q = (x, y, z, w)."""
    x, y, z, w = (float(t) for t in q)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy], [xy + wz, 1.0 - (xx + zz), yz - wx], [xz - wy, yz + wx, 1.0 - (xx + yy)]])


def make_transform(basis=None, origin=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Flattened btTransform: 9 basis doubles (row-major) + 3 origin doubles."""
    b = np.eye(3) if basis is None else np.asarray(basis, dtype=np.float64).reshape(3, 3)
    return np.concatenate([b.reshape(9), np.asarray(origin, dtype=np.float64).reshape(3)])


def transform_mul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """btTransform::operator* (btTransform.h:252-257): basis = A.b * B.b, origin = A(B.origin)."""
    Ab, Bb = A[:9].reshape(3, 3), B[:9].reshape(3, 3)
    out_b = np.empty((3, 3))
    for r in range(3):
        for c in range(3):  # btMatrix3x3 operator*: m[r].dot(column c) = tdotx/y/z, left-to-right sums
            out_b[r, c] = Ab[r, 0] * Bb[0, c] + Ab[r, 1] * Bb[1, c] + Ab[r, 2] * Bb[2, c]
    o = B[9:12]
    out_o = np.array([Ab[r, 0] * o[0] + Ab[r, 1] * o[1] + Ab[r, 2] * o[2] + A[9 + r] for r in range(3)])
    return np.concatenate([out_b.reshape(9), out_o])


def new_body_states(n: int) -> np.ndarray:
    s = np.zeros(n, dtype=BODY_STATE_DTYPE)
    eye = make_transform()
    s["T_CG"][:] = eye
    s["T_C"][:] = eye
    s["T_O_basis"][:] = np.eye(3).reshape(9)
    return s


# ------------------------------------------------------------------------------------------------ context
class SfhError(RuntimeError):
    pass


class HydroContext:
    """One GPU. Owns the sfh_ctx; meshes are uploaded once, `compute` is one recompute step."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = _abi.load_library()
        self.ctx = C.c_void_p()
        rc = (self.lib.sfh_create_on_stream(device, C.c_void_p(stream), C.byref(self.ctx)) if stream
              else self.lib.sfh_create(device, C.byref(self.ctx)))
        if rc != 0:
            raise SfhError(f"sfh_create({device}) failed ({rc}): {self.lib.sfh_last_error(None).decode()}")
        self.device = device
        self._waves_version = None
        self._keep = []  # buffers the library may still read asynchronously

    def close(self):
        if self.ctx:
            self.lib.sfh_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SfhError(f"{what} failed ({rc}): {self.lib.sfh_last_error(self.ctx).decode()}")

    # -- static geometry
    def upload_mesh(self, mesh: Mesh, keep_order: bool = False) -> int:
        mid = C.c_uint32()
        self._check(self.lib.sfh_mesh_upload(self.ctx, mesh.vertex_data.ctypes.data, mesh.stride_bytes, mesh.n_verts,
                                             mesh.faces.ctypes.data, mesh.n_faces, _abi.SFH_MESH_KEEP_ORDER if keep_order else 0,
                                             C.byref(mid)), "sfh_mesh_upload")
        return mid.value

    def mesh_stats(self, mesh_id: int) -> _abi.sfh_mesh_stats:
        st = _abi.sfh_mesh_stats()
        self._check(self.lib.sfh_mesh_get_stats(self.ctx, mesh_id, C.byref(st)), "sfh_mesh_get_stats")
        return st

    def release_mesh(self, mesh_id: int):
        self._check(self.lib.sfh_mesh_release(self.ctx, mesh_id), "sfh_mesh_release")

    # -- per step
    def _sync_waves(self, ocean: Ocean):
        if ocean.hasWaves() and not ocean.generated_waves and self._waves_version != ocean.waves_version:
            w = ocean.waves_struct()
            self._check(self.lib.sfh_waves_upload(self.ctx, C.byref(w)), "sfh_waves_upload")
            self._waves_version = ocean.waves_version

    # --- wave-field producer (row N3): OpenGLOcean::InitializeSimulation / Simulate on the device
    def wavegen_init(self, params: _abi.sfh_wave_params):
        self._check(self.lib.sfh_wavegen_init(self.ctx, C.byref(params)), "sfh_wavegen_init")
        self._wavegen_n = 1 << params.passes

    def wavegen_simulate(self, t: float):
        """h(k,t) + FFT at time t; the result becomes the context's wave field (as if its fftData had been uploaded)."""
        self._check(self.lib.sfh_wavegen_simulate(self.ctx, float(t)), "sfh_wavegen_simulate")
        self._waves_version = -1  # an uploaded field, if any, is no longer resident

    def wavegen_download(self) -> np.ndarray:
        n = self._wavegen_n
        out = np.zeros((4, n, n, 4), dtype=np.float32)
        self._check(self.lib.sfh_wavegen_download(self.ctx, out.ctypes.data, out.size), "sfh_wavegen_download")
        return out

    def wavegen_spectrum(self):
        n = self._wavegen_n
        s12, s34 = np.zeros(n * n * 4, dtype=np.float32), np.zeros(n * n * 4, dtype=np.float32)
        self._check(self.lib.sfh_wavegen_get_spectrum(self.ctx, s12.ctypes.data, s34.ctypes.data, s12.size), "sfh_wavegen_get_spectrum")
        return s12, s34

    def compute(self, ocean: Ocean, states: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """One recompute of all bodies: host buffers in, host buffers out (sfh_step)."""
        assert states.dtype == BODY_STATE_DTYPE and states.flags.c_contiguous
        self._sync_waves(ocean)
        if out is None:
            out = np.zeros(states.shape[0], dtype=INFO_DTYPE)
        o = ocean.to_struct()
        self._check(self.lib.sfh_step(self.ctx, C.byref(o), states.ctypes.data, states.shape[0], out.ctypes.data), "sfh_step")
        return out

    def compute_async(self, ocean_struct, states_ptr: int, n: int, out_ptr: int):
        self._check(self.lib.sfh_step_async(self.ctx, C.byref(ocean_struct), states_ptr, n, out_ptr), "sfh_step_async")

    def sync(self):
        self._check(self.lib.sfh_sync(self.ctx), "sfh_sync")

    def compute_debug(self, ocean: Ocean, states: np.ndarray, n_faces_total: int):
        """As compute, plus one SFH_FACE_* code per face (original face order, bodies concatenated)."""
        self._sync_waves(ocean)
        out = np.zeros(states.shape[0], dtype=INFO_DTYPE)
        codes = np.zeros(max(n_faces_total, 1), dtype=np.uint8)
        o = ocean.to_struct()
        self._check(self.lib.sfh_step_debug(self.ctx, C.byref(o), states.ctypes.data, states.shape[0], out.ctypes.data,
                                            codes.ctypes.data, codes.size), "sfh_step_debug")
        return out, codes[:n_faces_total]

    def compute_aero(self, atm: Atmosphere, states: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """One recompute of the air drag of all bodies (sfh_aero_step): Fda/Tda arrive in out['Fdq'] / out['Tdq']."""
        assert states.dtype == BODY_STATE_DTYPE and states.flags.c_contiguous
        if out is None:
            out = np.zeros(states.shape[0], dtype=INFO_DTYPE)
        a = atm.to_struct()
        self._check(self.lib.sfh_aero_step(self.ctx, C.byref(a), states.ctypes.data, states.shape[0], out.ctypes.data), "sfh_aero_step")
        return out

    def compute_aero_debug(self, atm: Atmosphere, states: np.ndarray, n_faces_total: int):
        out = np.zeros(states.shape[0], dtype=INFO_DTYPE)
        codes = np.zeros(max(n_faces_total, 1), dtype=np.uint8)
        a = atm.to_struct()
        self._check(self.lib.sfh_aero_step_debug(self.ctx, C.byref(a), states.ctypes.data, states.shape[0], out.ctypes.data,
                                                 codes.ctypes.data, codes.size), "sfh_aero_step_debug")
        return out, codes[:n_faces_total]

    def compute_aero_device(self, atm_struct, d_states_ptr: int, n: int, d_out_ptr: int):
        self._check(self.lib.sfh_aero_step_device(self.ctx, C.byref(atm_struct), d_states_ptr, n, d_out_ptr), "sfh_aero_step_device")

    def compute_cables(self, ocean: Ocean, cables, node_pos: np.ndarray, node_vel: np.ndarray) -> np.ndarray:
        """CableEntity::ComputeHydrodynamicForces for a batch of cables (sfh_cable_step). cables: iterable of
        (first_node, n_nodes, radius, flags). Returns (n_nodes_total, 9): Fb, Fdq, Fdf per node."""
        self._sync_waves(ocean)
        pos = np.ascontiguousarray(node_pos, dtype=np.float64).reshape(-1, 3)
        vel = np.ascontiguousarray(node_vel, dtype=np.float64).reshape(-1, 3)
        assert pos.shape == vel.shape
        cs = (_abi.sfh_cable * max(len(cables), 1))()
        for i, (first, n, radius, flags) in enumerate(cables):
            cs[i] = _abi.sfh_cable(int(first), int(n), int(flags), 0, float(radius))
        out = np.zeros((pos.shape[0], 9), dtype=np.float64)
        o = ocean.to_struct()
        self._check(self.lib.sfh_cable_step(self.ctx, C.byref(o), cs, len(cables), pos.ctypes.data, vel.ctypes.data, pos.shape[0],
                                            out.ctypes.data), "sfh_cable_step")
        return out

    def compute_device(self, ocean_struct, d_states_ptr: int, n: int, d_out_ptr: int, hints: int = 0):
        """Device-resident step (sfh_step_device / sfh_step_device_hinted): asynchronous on the context's stream."""
        self._check(self.lib.sfh_step_device_hinted(self.ctx, C.byref(ocean_struct), d_states_ptr, n, d_out_ptr, hints), "sfh_step_device_hinted")

    def GetDepth(self, ocean: Ocean, points: np.ndarray) -> np.ndarray:
        """Ocean::GetDepth(glm::vec3) for an [n,3] float32 array of world points."""
        self._sync_waves(ocean)
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(p.shape[0], dtype=np.float32)
        o = ocean.to_struct()
        self._check(self.lib.sfh_query_depth(self.ctx, C.byref(o), p.ctypes.data, p.shape[0], out.ctypes.data), "sfh_query_depth")
        return out

    def GetFluidVelocity(self, ocean: Ocean, points: np.ndarray) -> np.ndarray:
        """Ocean::GetFluidVelocity(glm::vec3) for an [n,3] float32 array of world points."""
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((p.shape[0], 3), dtype=np.float32)
        o = ocean.to_struct()
        self._check(self.lib.sfh_query_fluid_velocity(self.ctx, C.byref(o), p.ctypes.data, p.shape[0], out.ctypes.data),
                    "sfh_query_fluid_velocity")
        return out

    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.sfh_host_alloc(self.ctx, nbytes, C.byref(p)), "sfh_host_alloc")
        return p.value

    def host_free(self, ptr: int):
        self._check(self.lib.sfh_host_free(self.ctx, ptr), "sfh_host_free")

    def pinned_array(self, n: int, dtype) -> np.ndarray:
        """numpy view over pinned host memory owned by the library (freed with the context's lifetime)."""
        dt = np.dtype(dtype)
        ptr = self.host_alloc(max(n * dt.itemsize, 1))
        buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dt, count=n)
        self._keep.append((ptr, buf))
        return arr

    def set_profiling(self, on: bool):
        self._check(self.lib.sfh_set_profiling(self.ctx, 1 if on else 0), "sfh_set_profiling")

    def timings(self) -> _abi.sfh_timings:
        t = _abi.sfh_timings()
        self._check(self.lib.sfh_get_timings(self.ctx, C.byref(t)), "sfh_get_timings")
        return t


# ------------------------------------------------------------------------------------------------ compounds
def compound_part_states(compound: np.ndarray, parts: np.ndarray, external) -> np.ndarray:
    """Body states to submit for the parts of ONE compound (Compound.cpp:291-437): every part is evaluated with the
    compound's T_CG, velocities and AABB (so all parts resolve to the compound's BodyFluidPosition), its own T_C /
    T_O_basis / Cd / Cf / mesh, the full dispatcher (SFH_F_DISPATCH: reallisticBuoyancy &= part.isBuoyant(), per-part
    CorrectHydrodynamicForces), and no damping for internal parts."""
    st = parts.copy()
    c = compound
    for k in ("T_CG", "v", "omega", "aabb_min", "aabb_max"):
        st[k] = c[k]
    st["mode"] = c["mode"]
    base = int(c["flags"]) & (SFH_F_DAMPING | SFH_F_BUOYANCY)
    for i in range(len(st)):
        f = base | SFH_F_DISPATCH | (int(parts["flags"][i]) & SFH_F_BUOYANT)
        if not external[i]:
            f &= ~SFH_F_DAMPING
        st["flags"][i] = f
    return st


def sum_compound(part_infos: np.ndarray, parts: np.ndarray, external, compound: np.ndarray, ocean: Ocean) -> np.ndarray:
    """Host half of Compound::ComputeHydrodynamicForces: combine the per-part results of `compound_part_states` in part
    order exactly like Compound.cpp:315-435, including the temporaries that persist across parts (Fbp, Tbp, Vsubp keep
    the previous part's value when a part's buoyancy is not assigned) and the compound-level analytic buoyancy when INSIDE."""
    out = np.zeros((), dtype=INFO_DTYPE)
    pos = int(part_infos["position"][0]) if len(part_infos) else SFH_POS_OUTSIDE
    out["position"] = pos
    flags = int(compound["flags"])
    damping, rb = bool(flags & SFH_F_DAMPING), bool(flags & SFH_F_BUOYANCY)
    def count(r):  # validation counters of the parts the reference actually evaluates
        for k in ("class_count", "n_degenerate", "n_quadratic", "n_skin", "n_faces"):
            out[k] = out[k] + r[k]

    if pos == SFH_POS_OUTSIDE:
        out["wrote_buoyancy"] = out["wrote_damping"] = 1
        return out
    if pos == SFH_POS_INSIDE:
        if flags & SFH_F_BUOYANT:  # Compound.cpp:318-322, fp64 Bullet expressions
            s = -float(compound["volume"]) * ocean.density
            Fb = np.array([0.0 * s, 0.0 * s, ocean.gravity * s])
            T = compound["T_CG"]
            pcb, org = compound["P_CB"], T[9:12]
            wcb = np.array([T[3 * r] * pcb[0] + T[3 * r + 1] * pcb[1] + T[3 * r + 2] * pcb[2] + org[r] for r in range(3)])
            d = wcb - org
            out["Fb"] = Fb
            out["Tb"] = [d[1] * Fb[2] - d[2] * Fb[1], d[2] * Fb[0] - d[0] * Fb[2], d[0] * Fb[1] - d[1] * Fb[0]]
            out["wrote_buoyancy"] = 1
        if damping:
            out["wrote_damping"] = 1
            for i in range(len(part_infos)):
                if external[i]:
                    for k in ("Fdq", "Tdq", "Fdf", "Tdf"):
                        out[k] = out[k] + part_infos[k][i]
                    count(part_infos[i])
        out["Swet"], out["Vsub"] = compound["surface"], compound["volume"]
        return out
    if not (rb or damping):
        return out
    out["wrote_buoyancy"], out["wrote_damping"] = int(rb), int(damping)
    Fbp, Tbp, Vsubp = np.zeros(3), np.zeros(3), 0.0
    for i in range(len(part_infos)):
        r = part_infos[i]
        p_rb = rb and bool(int(parts["flags"][i]) & SFH_F_BUOYANT)
        if not external[i] and not p_rb:
            continue
        count(r)
        if r["wrote_buoyancy"]:
            Fbp, Tbp, Vsubp = r["Fb"].copy(), r["Tb"].copy(), float(r["Vsub"])
        out["Fb"] = out["Fb"] + Fbp
        out["Tb"] = out["Tb"] + Tbp
        out["Vsub"] = out["Vsub"] + Vsubp
        if external[i]:
            for k in ("Fdq", "Tdq", "Fdf", "Tdf"):
                out[k] = out[k] + r[k]
            out["Swet"] = out["Swet"] + r["Swet"]
    return out


def sum_compound_aero(part_infos: np.ndarray, external) -> np.ndarray:
    """Host half of Compound::ComputeAerodynamicForces (Compound.cpp:439-464): Fda/Tda of the external parts, each already
    corrected with its own proxy by sfh_aero_step (SFH_F_DISPATCH), summed in part order."""
    out = np.zeros(1, dtype=INFO_DTYPE)
    out["wrote_damping"] = 1
    for i in range(part_infos.shape[0]):
        if not external[i]:
            continue
        out["Fdq"][0] += part_infos["Fdq"][i]
        out["Tdq"][0] += part_infos["Tdq"][i]
        for k in ("n_faces", "n_degenerate", "n_quadratic"):
            out[k][0] += part_infos[k][i]
        out["class_count"][0][7] += part_infos["class_count"][i][7]
    return out
