"""Synthetic meshes and query sets of the shapes BASELINE.json / SURVEY.md §8(d) name.

numpy only.  Everything is float32 / int32, deterministic for a given seed, and cheap enough to
regenerate on every run (no cache files).  Used by tests/ and bench.py; not part of the query path.
"""

from __future__ import annotations

import numpy as np

# unit cube of the reference's golden tests (warp/tests/geometry/test_mesh.py:14-55)
CUBE_POINTS = np.array(
    [
        (0.5, -0.5, 0.5), (-0.5, -0.5, 0.5), (0.5, 0.5, 0.5), (-0.5, 0.5, 0.5),
        (-0.5, -0.5, -0.5), (0.5, -0.5, -0.5), (-0.5, 0.5, -0.5), (0.5, 0.5, -0.5),
    ],
    dtype=np.float32,
)  # fmt: skip
CUBE_INDICES_RH = np.array(
    [0, 3, 1, 0, 2, 3, 4, 7, 5, 4, 6, 7, 6, 2, 7, 6, 3, 2, 5, 1, 4, 5, 0, 1, 5, 2, 0, 5, 7, 2, 1, 6, 4, 1, 3, 6],
    dtype=np.int32,
)
CUBE_INDICES_LH = np.array(
    [0, 1, 3, 0, 3, 2, 4, 5, 7, 4, 7, 6, 6, 7, 2, 6, 2, 3, 5, 4, 1, 5, 1, 0, 5, 0, 2, 5, 2, 7, 1, 4, 6, 1, 6, 3],
    dtype=np.int32,
)


def icosphere(subdivisions: int, radius: float = 1.0):
    """Subdivided icosahedron: T = 20 * 4**subdivisions triangles, V = 10 * 4**subdivisions + 2."""
    t = (1.0 + 5.0**0.5) / 2.0
    verts = np.array(
        [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)],
        dtype=np.float64,
    )  # fmt: skip
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    faces = np.array(
        [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)],
        dtype=np.int64,
    )  # fmt: skip
    for _ in range(subdivisions):
        nv = verts.shape[0]
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        keys = e[:, 0] * nv + e[:, 1]
        uniq, inv = np.unique(keys, return_inverse=True)
        a, b = uniq // nv, uniq % nv
        mid = verts[a] + verts[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        verts = np.concatenate([verts, mid], axis=0)
        nf = faces.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf : 2 * nf], nv + inv[2 * nf :]
        v0, v1, v2 = faces[:, 0], faces[:, 1], faces[:, 2]
        faces = np.concatenate(
            [np.stack([v0, m01, m20], 1), np.stack([v1, m12, m01], 1), np.stack([v2, m20, m12], 1),
             np.stack([m01, m12, m20], 1)],
            axis=0,
        )  # fmt: skip
    return (verts * radius).astype(np.float32), faces.astype(np.int32).reshape(-1)


def noisy_sphere(subdivisions: int = 8, noise: float = 0.02, seed: int = 1):
    """Icosphere with vertex radius 1 + noise * N(0,1) (config C2: 8 subdivisions = 1 310 720 tris)."""
    p, idx = icosphere(subdivisions)
    rng = np.random.default_rng(seed)
    r = (1.0 + noise * rng.standard_normal(p.shape[0])).astype(np.float32)
    return (p * r[:, None]).astype(np.float32), idx


def renoise_sphere(points, noise: float = 0.02, seed: int = 3):
    """New radial noise on the same topology (the refit input of C2)."""
    rng = np.random.default_rng(seed)
    unit = points / np.linalg.norm(points.astype(np.float64), axis=1, keepdims=True)
    r = 1.0 + noise * rng.standard_normal(points.shape[0])
    return (unit * r[:, None]).astype(np.float32)


def _grid_indices(nx: int, ny: int):
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="ij")
    v00 = (i * ny + j).reshape(-1)
    v10, v01, v11 = v00 + ny, v00 + 1, v00 + ny + 1
    tris = np.stack([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=1)  # 2 per cell
    return tris.reshape(-1).astype(np.int32)


def heightfield(n: int = 2237, seed: int = 4):
    """n x n vertex grid on [0,1]^2 with 6 octaves of sin*cos relief (C3: n=2237 -> 9 999 392 tris)."""
    rng = np.random.default_rng(seed)
    phi, psi = rng.uniform(0, 2 * np.pi, 6), rng.uniform(0, 2 * np.pi, 6)
    x = np.linspace(0.0, 1.0, n)
    xx, yy = np.meshgrid(x, x, indexing="ij")
    z = np.zeros_like(xx)
    for k in range(6):
        z += 0.25 * 2.0**-k * np.sin(2.0**k * 7.0 * xx + phi[k]) * np.cos(2.0**k * 5.0 * yy + psi[k])
    pts = np.stack([xx, yy, z], axis=-1).reshape(-1, 3).astype(np.float32)
    return pts, _grid_indices(n, n)


def cloth(n: int = 1415, frame: int = 0):
    """n x n cloth grid on [0,1]^2, z = 0.05 sin(12x + 0.05f) cos(9y + 0.03f) (C4: n=1415)."""
    x = np.linspace(0.0, 1.0, n)
    xx, yy = np.meshgrid(x, x, indexing="ij")
    z = 0.05 * np.sin(12.0 * xx + 0.05 * frame) * np.cos(9.0 * yy + 0.03 * frame)
    pts = np.stack([xx, yy, z], axis=-1).reshape(-1, 3).astype(np.float32)
    return pts, _grid_indices(n, n)


def box_queries(points, count: int, scale: float = 1.2, seed: int = 2):
    """Uniform points in the mesh AABB scaled by ``scale`` about its centre."""
    lo, hi = points.min(axis=0).astype(np.float64), points.max(axis=0).astype(np.float64)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo) * scale
    rng = np.random.default_rng(seed)
    return (c + (rng.random((count, 3)) * 2.0 - 1.0) * h).astype(np.float32)


def cube_queries(count: int, half: float = 1.5, seed: int = 42):
    """Uniform points in [-half, half]^3 (config C1)."""
    rng = np.random.default_rng(seed)
    return ((rng.random((count, 3)) * 2.0 - 1.0) * half).astype(np.float32)


def pinhole_rays(width: int, height: int, eye=(0.5, -0.6, 0.9), look_at=(0.5, 0.5, 0.0), vfov_deg: float = 50.0):
    """Row-major primary rays of a pinhole camera (config C3: 4096 x 4096)."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(look_at, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, (0.0, 0.0, 1.0))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    th = np.tan(np.radians(vfov_deg) * 0.5)
    aspect = width / height
    px = ((np.arange(width) + 0.5) / width * 2.0 - 1.0) * th * aspect
    py = (1.0 - (np.arange(height) + 0.5) / height * 2.0) * th
    d = fwd[None, None, :] + px[None, :, None] * right[None, None, :] + py[:, None, None] * up[None, None, :]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    starts = np.broadcast_to(eye.astype(np.float32), (height * width, 3)).copy()
    return starts, d.reshape(-1, 3).astype(np.float32)


def random_rays(points, count: int, seed: int = 7):
    """Rays from a shell around the mesh aimed at random points inside its AABB (parity tests)."""
    rng = np.random.default_rng(seed)
    lo, hi = points.min(axis=0).astype(np.float64), points.max(axis=0).astype(np.float64)
    c, r = 0.5 * (lo + hi), 0.5 * np.linalg.norm(hi - lo)
    o = rng.standard_normal((count, 3))
    o = c + 1.5 * r * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = lo + rng.random((count, 3)) * (hi - lo)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)
