"""Upload-time mesh re-layout (stonefish_b200/csrc/mesh_chunker.hpp): every face appears exactly once, local indices
decode to the original vertex positions bit-for-bit, chunk sizes respect the 256/256 limits and the 16-byte bulk-copy
granularity, and next_bytes chains the chunks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from stonefish_b200 import meshes
from stonefish_b200.hydro import Mesh

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def chk():
    so = os.path.join(HERE, "_build", "libchunker_test.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "chunker_wrap.cpp")])
    lib = C.CDLL(so)
    lib.chk_build.restype = C.c_void_p
    lib.chk_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int]
    for n, r in (("chk_n_chunks", C.c_uint32), ("chk_bytes", C.c_uint64), ("chk_data", C.c_void_p), ("chk_off", C.c_void_p),
                 ("chk_orig", C.c_void_p), ("chk_chunk_verts", C.c_uint32)):
        getattr(lib, n).restype = r
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.chk_free.argtypes = [C.c_void_p]
    return lib


def _decode(lib, h):
    n = lib.chk_n_chunks(h)
    data = np.ctypeslib.as_array(C.cast(lib.chk_data(h), C.POINTER(C.c_uint8)), shape=(lib.chk_bytes(h),)).copy()
    off = np.ctypeslib.as_array(C.cast(lib.chk_off(h), C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
    orig = np.ctypeslib.as_array(C.cast(lib.chk_orig(h), C.POINTER(C.c_uint32)), shape=(n * 256,)).copy()
    return n, data, off, orig


def _check_roundtrip(chk, m, keep, finite=True):
    h = chk.chk_build(m.vertex_data.ctypes.data, m.stride_bytes, m.n_verts, m.faces.ctypes.data, m.n_faces, keep)
    assert h
    try:
        n, data, off, orig = _decode(chk, h)
        assert off[0] == 0 and off[-1] * 16 == data.size
        seen = np.zeros(m.n_faces, dtype=np.int64)
        total_v = 0
        for c in range(n):
            base = int(off[c]) * 16
            nv, nf, next_bytes, first_slot = (int(t) for t in np.frombuffer(data, np.uint32, 4, base))
            nv4, nf4 = (nv + 3) & ~3, (nf + 3) & ~3
            size = 32 + 12 * nv4 + 4 * nf4
            assert 1 <= nf <= 256 and 1 <= nv <= 256 and first_slot == c * 256
            assert size == (int(off[c + 1]) - int(off[c])) * 16 and size % 16 == 0 and size <= 4128
            assert next_bytes == ((int(off[c + 2]) - int(off[c + 1])) * 16 if c + 1 < n else 0)
            x = np.frombuffer(data, np.float32, nv, base + 32)
            y = np.frombuffer(data, np.float32, nv, base + 32 + 4 * nv4)
            z = np.frombuffer(data, np.float32, nv, base + 32 + 8 * nv4)
            f = np.frombuffer(data, np.uint32, nf, base + 32 + 12 * nv4)
            sx, sy, sz, sr = (float(t) for t in np.frombuffer(data, np.float32, 4, base + 16))  # bounding sphere of the vertices
            with np.errstate(invalid="ignore"):
                dist = np.sqrt((x.astype(np.float64) - sx) ** 2 + (y.astype(np.float64) - sy) ** 2 + (z.astype(np.float64) - sz) ** 2)
            if finite:
                assert dist.max() <= sr <= dist.max() * 1.001 + 1e-6
            loc = np.stack([f & 255, (f >> 8) & 255, (f >> 16) & 255], axis=1)
            assert (loc < nv).all() and (f >> 24 == 0).all()
            faces = orig[c * 256: c * 256 + nf]
            assert (orig[c * 256 + nf: (c + 1) * 256] == 0xFFFFFFFF).all()
            seen[faces] += 1
            want = m.positions[m.faces[faces]]                      # [nf, 3 vertices, xyz]
            got = np.stack([x[loc], y[loc], z[loc]], axis=2)
            assert got.tobytes() == want.tobytes()
            total_v += nv
        assert (seen == 1).all()
        assert total_v == chk.chk_chunk_verts(h)
        if keep:  # original order is preserved across chunks
            assert (np.diff(orig[orig != 0xFFFFFFFF].astype(np.int64)) > 0).all()
    finally:
        chk.chk_free(h)


@pytest.mark.parametrize("keep", [0, 1])
@pytest.mark.parametrize("mesh", ["ellipsoid", "icosphere_strided", "degenerate_box", "soup"])
def test_chunks_reproduce_the_mesh(chk, mesh, keep):
    if mesh == "ellipsoid":
        m = meshes.uv_ellipsoid(2.5, 0.3, 0.3, 100, 51)
    elif mesh == "icosphere_strided":
        m = meshes.icosphere(0.6, 3, stride_floats=11)
    elif mesh == "degenerate_box":
        b = meshes.box(1.0, 0.5, 0.2, 2)
        m = Mesh(b.positions.copy(), np.concatenate([b.faces, [[0, 0, 1], [2, 2, 2]]]).astype(np.uint32))
    else:  # unshared vertices in random order: worst case for locality, 3 fresh vertices per face
        r = meshes.splitmix_uniform(9, 900 * 3).reshape(-1, 3).astype(np.float32)
        m = Mesh(r, np.arange(900, dtype=np.uint32).reshape(-1, 3))
    _check_roundtrip(chk, m, keep)


def test_morton_order_improves_vertex_sharing(chk):
    m = meshes.uv_ellipsoid(1.2, 0.15, 0.15, 250, 101)
    hs = [chk.chk_build(m.vertex_data.ctypes.data, m.stride_bytes, m.n_verts, m.faces.ctypes.data, m.n_faces, k) for k in (0, 1)]
    sorted_v, kept_v = (chk.chk_chunk_verts(h) for h in hs)
    for h in hs:
        chk.chk_free(h)
    assert sorted_v < 0.85 * kept_v  # strips of a UV grid share far fewer vertices per 256 faces than Morton patches
    assert sorted_v < 0.9 * m.n_faces


def test_bad_index_is_rejected(chk):
    v = np.zeros((3, 3), np.float32)
    f = np.array([[0, 1, 3]], np.uint32)
    assert not chk.chk_build(v.ctypes.data, 12, 3, f.ctypes.data, 1, 0)


def test_random_meshes_roundtrip(chk):
    """Property test over ragged random meshes: any vertex count, any face count (also not a multiple of anything), faces that
    repeat a vertex, vertices nobody uses, clustered and scattered index patterns, strided vertex records and non-finite
    coordinates: every face once, bits preserved, limits kept, in both orders."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 700), st.integers(1, 1500), st.integers(0, 2 ** 31), st.sampled_from([3, 6, 11]), st.integers(0, 3), st.integers(0, 1))
    def run(n_verts, n_faces, seed, stride_floats, pattern, keep):
        rng = np.random.default_rng(seed)
        pos = (rng.standard_normal((n_verts, 3)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
        finite = True
        if pattern == 3 and n_verts > 2:  # NaN / Inf coordinates travel bit for bit (the kernels let them propagate like the reference)
            pos[rng.integers(0, n_verts)] = [np.nan, np.inf, -np.inf]
            finite = False
        if pattern == 0:    # scattered: almost no sharing
            faces = rng.integers(0, n_verts, (n_faces, 3))
        elif pattern == 1:  # clustered: consecutive faces share vertices
            base = rng.integers(0, n_verts, n_faces)
            faces = (base[:, None] + rng.integers(0, 4, (n_faces, 3))) % n_verts
        else:               # strip-like, with repeated vertices inside a face now and then
            k = np.arange(n_faces)
            faces = np.stack([k % n_verts, (k + 1) % n_verts, (k + 1 + (k % 7 != 0)) % n_verts], axis=1)
        m = Mesh(pos, faces.astype(np.uint32), vertex_stride_floats=stride_floats)
        _check_roundtrip(chk, m, keep, finite)

    run()
