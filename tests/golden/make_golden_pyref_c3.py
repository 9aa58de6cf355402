"""Generates tests/golden/c3_4k_sample_pyref.npz: 400 pixels of config 3 (Suzanne subdivided x5 = 991 232 faces, 3840x2160,
lights[0..1]) from the independent restatement of make_golden_pyref.py (own scene reader, own math.Tan, tests/pyref.py
arithmetic, numpy candidate search over all 991 232 faces) AND an independent midpoint subdivision written here from the
rule SURVEY s8d states: each triangle splits 1 -> 4 at its edge midpoints, 5 rounds; face f becomes faces 4f..4f+3 (the corner
triangles of v0, v1, v2, then the centre triangle); a new vertex is the float64 mean of the edge's endpoints stored as float32,
a new normal the normalised mean of the endpoint normals.  Nothing of the oracle or of the product is imported.

Run (from the repo root, ~3 min):  python tests/golden/make_golden_pyref_c3.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_golden_pyref as mg  # noqa: E402
import pyref  # noqa: E402


def subdivide_once(tris_v, tris_n):
    """tris_v, tris_n: [F, 3, 3] per-corner positions / unit normals -> [4F, 3, 3] each."""
    def mid_v(a, b):
        return (0.5 * (a + b)).astype(np.float32).astype(np.float64)

    def mid_n(a, b):
        m = 0.5 * (a + b)
        mag = np.sqrt(m[:, 0] * m[:, 0] + m[:, 1] * m[:, 1] + m[:, 2] * m[:, 2])
        return m / mag[:, None]
    v0, v1, v2 = tris_v[:, 0], tris_v[:, 1], tris_v[:, 2]
    n0, n1, n2 = tris_n[:, 0], tris_n[:, 1], tris_n[:, 2]
    # an edge midpoint must not depend on the direction the edge is walked in: order the endpoints canonically
    def ordered(a, b, na, nb):
        key_a = a[:, 0] * 1e6 + a[:, 1] * 1e3 + a[:, 2]
        key_b = b[:, 0] * 1e6 + b[:, 1] * 1e3 + b[:, 2]
        sw = (key_b < key_a)[:, None]
        return np.where(sw, b, a), np.where(sw, a, b), np.where(sw, nb, na), np.where(sw, na, nb)
    a, b, na, nb = ordered(v0, v1, n0, n1); m01, k01 = mid_v(a, b), mid_n(na, nb)
    a, b, na, nb = ordered(v1, v2, n1, n2); m12, k12 = mid_v(a, b), mid_n(na, nb)
    a, b, na, nb = ordered(v2, v0, n2, n0); m20, k20 = mid_v(a, b), mid_n(na, nb)
    F = tris_v.shape[0]
    ov, on = np.empty((F, 4, 3, 3)), np.empty((F, 4, 3, 3))
    ov[:, 0] = np.stack([v0, m01, m20], axis=1); on[:, 0] = np.stack([n0, k01, k20], axis=1)
    ov[:, 1] = np.stack([m01, v1, m12], axis=1); on[:, 1] = np.stack([k01, n1, k12], axis=1)
    ov[:, 2] = np.stack([m20, m12, v2], axis=1); on[:, 2] = np.stack([k20, k12, n2], axis=1)
    ov[:, 3] = np.stack([m01, m12, m20], axis=1); on[:, 3] = np.stack([k01, k12, k20], axis=1)
    return ov.reshape(-1, 3, 3), on.reshape(-1, 3, 3)


def c3_scene():
    sc = mg.read_obj_scene(os.path.join(HERE, "example", "scene.json"))
    m = sc["meshes"][0]
    v, n, f = np.array(m["verts"]), np.array(m["normals"]), np.array(m["faces"])
    tv, tn = v[f[:, 0:3]], n[f[:, 3:6]]
    for _ in range(5):
        tv, tn = subdivide_once(tv, tn)
    F = tv.shape[0]
    idx = np.arange(3 * F).reshape(F, 3)
    faces = np.concatenate([idx, idx, np.zeros((F, 1), dtype=np.int64)], axis=1)
    sc["meshes"] = [{"verts": tv.reshape(-1, 3), "normals": tn.reshape(-1, 3), "faces": faces, "materials": m["materials"]}]
    sc["lights"] = sc["lights"][:2]
    return sc


class BigScene(mg.FastScene):
    """pyref.Scene keeps Python tuples per vertex: too heavy for 3 M corners.  The candidate search stays in numpy; the scalar
    pyref arithmetic only ever touches the candidate faces, whose corners are fetched on demand."""

    def __init__(self, sc):
        self.cam_pos, self.fwd, self.left, self.up = (tuple(float(c) for c in sc["camera"][k:k + 3]) for k in (0, 3, 6, 9))
        self.fov = float(sc["camera"][12])
        self.lights = [(tuple(float(c) for c in l[0:3]), tuple(float(c) for c in l[3:6])) for l in sc["lights"]]
        self.objs = [(tuple(float(c) for c in p), int(k)) for p, k in zip(sc["obj_pos"], sc["obj_mesh"])]
        m = sc["meshes"][0]
        self.V, self.N = np.asarray(m["verts"], dtype=np.float64), np.asarray(m["normals"], dtype=np.float64)
        self.F = np.asarray(m["faces"], dtype=np.int64)
        self.mats = [tuple(float(c) for c in r) for r in m["materials"]]
        P1 = self.V[self.F[:, 0]]
        self.np_meshes = [(P1, self.V[self.F[:, 1]] - P1, self.V[self.F[:, 2]] - P1)]
        self.meshes = [self]                              # object_intersection() is handed `mesh` = self

    def object_intersection(self, pos, mesh, o, d):
        ol = pyref.sub(o, pos)
        best = None
        for fi in self.candidates(0, ol, d):
            f = self.F[fi]
            P1, P2, P3 = (tuple(float(c) for c in self.V[f[k]]) for k in range(3))
            hit = pyref.tri_intersection(P1, P2, P3, ol, d)
            assert hit is not None
            ip, (r1, r2, r3) = hit
            N1, N2, N3 = (tuple(float(c) for c in self.N[f[3 + k]]) for k in range(3))
            normal = pyref.norm(pyref.add(pyref.add(pyref.scale(N1, r1), pyref.scale(N2, r2)), pyref.scale(N3, r3)))
            dist = pyref.length(pyref.sub(ol, ip))
            if best is None or dist < best[0]:
                best = (dist, ip, normal, self.mats[f[6]], int(fi))
        if best is None:
            return None
        return pyref.add(best[1], pos), best[2], best[3], best[4]


if __name__ == "__main__":
    sc = c3_scene()
    assert sc["meshes"][0]["faces"].shape[0] == 991232
    S = BigScene(sc)
    tan_half = mg.go_tan(float(sc["camera"][12]) / 2.0)
    W, H = 3840, 2160
    # 400 pixels: a coarse lattice over the part of the screen the object covers (most of them hit) + a few sure misses
    px = [(i, j) for i in range(1270, 2580, 66) for j in range(650, 1520, 44)]
    px += [(5, 7), (3835, 2150), (1920, 100), (100, 1080)]
    face, lit, col, rgb, sdl = mg.render(S, W, H, tan_half, px)
    np.savez_compressed(os.path.join(HERE, "c3_4k_sample_pyref.npz"), pixel_i=np.array([p[0] for p in px], dtype=np.int32),
                        pixel_j=np.array([p[1] for p in px], dtype=np.int32), face=face.astype(np.int32), colour_bits=col.view(np.uint64),
                        rgb=rgb, sdl=sdl, lit=lit.astype(np.uint8))
    print("C3 sample: %d pixels, %d hits" % (len(px), (face >= 0).sum()))
