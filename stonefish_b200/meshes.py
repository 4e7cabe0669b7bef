"""Deterministic synthetic physics meshes (stand-ins for SolidEntity::phyMesh) and a minimal OBJ reader.

Shapes follow what the reference's solids build (Library/src/entities/solids/*.cpp via OpenGLContent): a
subdivided icosphere (Sphere.cpp:63, 20*4^3 = 1280 faces), a box tessellated 2^3 x 2^3 per side (Box.cpp:72,
768 faces) and closed UV-ellipsoid hulls for the BASELINE configs. All faces are wound so that
cross(p2-p1, p3-p1) points out of the body, as the integrators assume.
"""
from __future__ import annotations

import numpy as np

from .hydro import Mesh


def uv_ellipsoid(a: float, b: float, c: float, segments: int, stacks: int, stride_floats: int = 3) -> Mesh:
    """Closed ellipsoid, semi-axes (a,b,c): F = 2*S*(R-1) faces, V = S*(R-1)+2 vertices."""
    S, R = int(segments), int(stacks)
    th = np.pi * np.arange(1, R, dtype=np.float64) / R           # polar angle of the R-1 rings
    ph = 2.0 * np.pi * np.arange(S, dtype=np.float64) / S
    ring = np.stack([a * np.cos(th)[:, None] * np.ones(S)[None, :],
                     b * np.sin(th)[:, None] * np.cos(ph)[None, :],
                     c * np.sin(th)[:, None] * np.sin(ph)[None, :]], axis=-1).reshape(-1, 3)
    verts = np.concatenate([[[a, 0.0, 0.0]], ring, [[-a, 0.0, 0.0]]]).astype(np.float32)
    top, bottom = 0, verts.shape[0] - 1

    def vid(r, s):
        return 1 + r * S + (s % S)

    s = np.arange(S)
    faces = [np.stack([np.full(S, top), vid(0, s), vid(0, s + 1)], axis=1)]
    for r in range(R - 2):
        v00, v01, v10, v11 = vid(r, s), vid(r, s + 1), vid(r + 1, s), vid(r + 1, s + 1)
        faces.append(np.stack([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)], axis=1).reshape(-1, 3))
    faces.append(np.stack([np.full(S, bottom), vid(R - 2, s + 1), vid(R - 2, s)], axis=1))
    return Mesh(verts, np.concatenate(faces).astype(np.uint32), stride_floats)


def icosphere(radius: float, subdivisions: int = 3, stride_floats: int = 3) -> Mesh:
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1),
         (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdivisions):
        cache: dict[tuple[int, int], int] = {}

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = verts[i] + verts[j]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for (i, j, k) in faces:
            a, b, c = mid(i, j), mid(j, k), mid(k, i)
            nf += [(i, a, c), (j, b, a), (k, c, b), (a, b, c)]
        faces = nf
    return Mesh((np.array(verts) * radius).astype(np.float32), np.array(faces, dtype=np.uint32), stride_floats)


def box(lx: float, ly: float, lz: float, subdivisions: int = 3, stride_floats: int = 3) -> Mesh:
    """Axis-aligned box of dimensions (lx,ly,lz), each side split into (2^n)^2 quads: 12*4^n faces."""
    n = 2 ** subdivisions
    h = np.array([lx, ly, lz], dtype=np.float64) / 2.0
    verts: list[np.ndarray] = []
    faces: list[tuple[int, int, int]] = []
    index: dict[tuple[int, int, int], int] = {}

    def vid(p):
        key = tuple(int(round(q)) for q in p)
        if key not in index:
            index[key] = len(verts)
            verts.append((np.array(key, dtype=np.float64) / n * 2.0 - 1.0) * h)
        return index[key]

    for axis in range(3):
        u, w = (axis + 1) % 3, (axis + 2) % 3
        for side in (0, n):
            for i in range(n):
                for j in range(n):
                    def P(di, dj):
                        p = [0, 0, 0]
                        p[axis], p[u], p[w] = side, i + di, j + dj
                        return vid(p)
                    q = [P(0, 0), P(1, 0), P(1, 1), P(0, 1)]
                    if side == 0:
                        q = q[::-1]
                    faces += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
    return Mesh(np.array(verts, dtype=np.float32), np.array(faces, dtype=np.uint32), stride_floats)


def jitter(mesh: Mesh, seed: int, amplitude: float = 1e-3) -> Mesh:
    """Copy of `mesh` with every vertex moved by U[-amplitude, amplitude]^3 (unique-mesh-per-instance runs)."""
    r = splitmix_uniform(seed, mesh.n_verts * 3).reshape(-1, 3)
    pos = mesh.positions.astype(np.float64) + (2.0 * r - 1.0) * amplitude
    return Mesh(pos.astype(np.float32), mesh.faces, mesh.vertex_data.shape[1])


def signed_volume(mesh: Mesh) -> float:
    p = mesh.positions.astype(np.float64)
    a, b, c = p[mesh.faces[:, 0]], p[mesh.faces[:, 1]], p[mesh.faces[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def surface_area(mesh: Mesh) -> float:
    p = mesh.positions.astype(np.float64)
    a, b, c = p[mesh.faces[:, 0]], p[mesh.faces[:, 1]], p[mesh.faces[:, 2]]
    return float(np.linalg.norm(np.cross(b - a, c - a), axis=1).sum() / 2.0)


def load_obj(path: str, stride_floats: int = 3) -> Mesh:
    """Positions and triangular faces of a Wavefront OBJ (v / f records only; polygons are fanned)."""
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                verts.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                ids = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                ids = [i - 1 if i > 0 else len(verts) + i for i in ids]
                for k in range(1, len(ids) - 1):
                    faces.append((ids[0], ids[k], ids[k + 1]))
    return Mesh(np.array(verts, dtype=np.float32), np.array(faces, dtype=np.uint32), stride_floats)


# ------------------------------------------------------------------------------------------------ RNG
def splitmix_uniform(seed: int, n: int) -> np.ndarray:
    """n doubles in [0,1) from SplitMix64 (vectorised, independent of numpy's generators)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
