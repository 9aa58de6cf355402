"""Deterministic synthetic workloads for BASELINE.json's configs (bench + parity tests).

C3: Suzanne with every triangle split 1 -> 4 at edge midpoints, 5 rounds (968 * 4^5 = 991 232 faces);
    a new vertex is the float64 mean of the edge's endpoints stored as float32 (what an OBJ export
    re-read through gwob would hold, mesh.go:176), a new normal the normalised mean.
C4: 10 M random triangles: centre ~ U([-1.5,1.5]^3), vertices = centre + U([-0.01,0.01]^3),
    counter-based splitmix64 stream, seed 0x5EED, stored as float32; no vertex normals (flat-normal
    path, object.go:93-95); default material (mesh.go:151); 8 white lights on a circle r=10 at z=10.
Meshes use the flat layout of include/rt_b200.h (verts, normals, faces[n,7], materials[n,10]).
"""
import numpy as np


def _normalise_rows(n):
    mag = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
    mag = np.where(mag == 0.0, 1.0, mag)
    return n / mag[:, None]


def subdivide(mesh, rounds):
    """Midpoint subdivision; keeps face order stable: face f -> faces 4f..4f+3
    (corner triangles of v0, v1, v2, then the centre triangle)."""
    verts = np.asarray(mesh["verts"], dtype=np.float64)
    norms = np.asarray(mesh["normals"], dtype=np.float64).reshape(-1, 3)
    faces = np.asarray(mesh["faces"], dtype=np.uint32).reshape(-1, 7)
    has_n = norms.shape[0] > 0
    for _ in range(rounds):
        nf = faces.shape[0]
        v = faces[:, 0:3].astype(np.int64)
        n = faces[:, 3:6].astype(np.int64)
        mat = faces[:, 6]
        # per-corner records (vertex index, normal index); an edge midpoint is keyed on the unordered pair of records
        nv = verts.shape[0]
        nn = max(norms.shape[0], 1)
        rec = v * nn + n                                     # [nf,3] unique id of (vertex, normal)
        e_a = np.stack([rec[:, 0], rec[:, 1], rec[:, 2]], axis=1)
        e_b = np.stack([rec[:, 1], rec[:, 2], rec[:, 0]], axis=1)
        lo = np.minimum(e_a, e_b).reshape(-1)
        hi = np.maximum(e_a, e_b).reshape(-1)
        key = lo * (nv * nn) + hi
        uniq, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
        ulo, uhi = lo[first_idx], hi[first_idx]
        mv = (0.5 * (verts[ulo // nn] + verts[uhi // nn])).astype(np.float32).astype(np.float64)
        mid_v = nv + np.arange(uniq.shape[0], dtype=np.int64)
        verts = np.concatenate([verts, mv], axis=0)
        if has_n:
            mn = _normalise_rows(0.5 * (norms[ulo % nn] + norms[uhi % nn]))
            mid_n = norms.shape[0] + np.arange(uniq.shape[0], dtype=np.int64)
            norms = np.concatenate([norms, mn], axis=0)
        else:
            mid_n = np.zeros(uniq.shape[0], dtype=np.int64)
        inv = inv.reshape(nf, 3)
        m01v, m12v, m20v = mid_v[inv[:, 0]], mid_v[inv[:, 1]], mid_v[inv[:, 2]]
        m01n, m12n, m20n = mid_n[inv[:, 0]], mid_n[inv[:, 1]], mid_n[inv[:, 2]]
        out = np.empty((nf, 4, 7), dtype=np.int64)
        out[:, 0] = np.stack([v[:, 0], m01v, m20v, n[:, 0], m01n, m20n, mat], axis=1)
        out[:, 1] = np.stack([m01v, v[:, 1], m12v, m01n, n[:, 1], m12n, mat], axis=1)
        out[:, 2] = np.stack([m20v, m12v, v[:, 2], m20n, m12n, n[:, 2], mat], axis=1)
        out[:, 3] = np.stack([m01v, m12v, m20v, m01n, m12n, m20n, mat], axis=1)
        faces = out.reshape(-1, 7).astype(np.uint32)
    return {"verts": verts, "normals": norms if has_n else np.zeros((0, 3)), "faces": faces,
            "materials": np.asarray(mesh["materials"], dtype=np.float64)}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform01(seed, counter):
    with np.errstate(over="ignore"):
        bits = _splitmix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + counter.astype(np.uint64))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


DEFAULT_MATERIAL = [0x10 / 255.0] * 3 + [1.0] * 3 + [0.0] * 3 + [0.0]    # mesh.go:151


def triangle_soup(n_tris, seed=0x5EED, extent=1.5, half=0.01):
    """C4 workload: n_tris random small triangles, no normals, default material."""
    idx = np.arange(n_tris, dtype=np.uint64)
    cen = np.stack([_uniform01(seed, idx * np.uint64(12) + np.uint64(k)) for k in range(3)], axis=1)
    cen = (cen * 2.0 - 1.0) * extent
    verts = np.empty((n_tris, 3, 3), dtype=np.float64)
    for v in range(3):
        off = np.stack([_uniform01(seed, idx * np.uint64(12) + np.uint64(3 + 3 * v + k)) for k in range(3)], axis=1)
        verts[:, v, :] = cen + (off * 2.0 - 1.0) * half
    verts = verts.reshape(-1, 3).astype(np.float32).astype(np.float64)
    faces = np.zeros((n_tris, 7), dtype=np.uint32)
    faces[:, 0:3] = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    return {"verts": verts, "normals": np.zeros((0, 3)), "faces": faces,
            "materials": np.asarray([DEFAULT_MATERIAL], dtype=np.float64)}


def ring_lights(n, radius=10.0, z=10.0):
    """n white point lights on a circle of `radius` in the plane z (C4)."""
    a = 2.0 * np.pi * np.arange(n) / n
    out = np.zeros((n, 6), dtype=np.float64)
    out[:, 0] = radius * np.cos(a)
    out[:, 1] = radius * np.sin(a)
    out[:, 2] = z
    out[:, 3:6] = 1.0
    return out


def orbit_camera(centre, radius, angle_rad, height, fov):
    """C5: camera on a circle of `radius` around `centre` looking at it; basis per camera.go:35-44."""
    pos = np.array([centre[0] + radius * np.sin(angle_rad), centre[1] + height, centre[2] + radius * np.cos(angle_rad)])
    d = np.asarray(centre, dtype=np.float64) - pos

    def norm(v):
        m = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        return np.array([v[0] / m, v[1] / m, v[2] / m])

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    forward = norm(d)
    left = norm(cross(d, np.array([0.0, 1.0, 0.0])))
    up = cross(left, forward)
    return np.concatenate([pos, forward, left, up, [fov]])
