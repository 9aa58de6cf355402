"""Generates the golden vectors tests/golden/c1_320x240_pyref.npz (the WHOLE C1 frame) and c2_1080p_sample_pyref.npz (a
1/97 pixel sample of C2) from a restatement that shares NOTHING with the oracle or the product:

  * its own reader of example/scene.json + suzanne.obj/.mtl (below), written from environment.go:162-234, mesh.go:109-213 and
    the published behaviour of gwob (float32 parse then widened, quad -> (0,1,2) + (2,3,0), faces in file order);
  * its own math.Tan (Cephes, as in Go's math/tan.go) for the frame constant;
  * the per-pixel arithmetic of tests/pyref.py (scalar Python floats, written from the Go sources) for the shading, and a
    numpy restatement of triangle.go:37-77 over all faces at once for the candidate search.  numpy's float64 ufuncs are
    single IEEE operations applied in the order written (no fusion), and the generator re-checks that search against the
    scalar pyref code on a few hundred rays before it writes anything.

The Go reference cannot run here (no toolchain), so this is still not reference OUTPUT; what it removes is the circularity
of a golden made by the oracle itself: GPU == golden and oracle == golden are now checks against a third party.

Run (from the repo root, ~2 min):  python tests/golden/make_golden_pyref.py
"""
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pyref  # noqa: E402


# ---- Go math.Tan (math/tan.go, Cephes) -----------------------------------------------------------------------------------
def go_tan(x):
    PI4A, PI4B, PI4C = 7.85398125648498535156e-1, 3.77489470793079817668e-8, 2.69515142907905952645e-15
    P = (-1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7)
    Q = (1.0, 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7, -5.38695755929454629881e7)
    if x == 0 or math.isnan(x):
        return x
    if math.isinf(x):
        return math.nan
    sign = False
    if x < 0:
        x, sign = -x, True
    assert x < (1 << 29)                      # the Payne-Hanek branch of newer Go versions is not needed for a field of view
    j = int(x * (4 / math.pi))
    y = float(j)
    if j & 1 == 1:
        j += 1
        y += 1
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    zz = z * z
    if zz > 1e-14:
        y = z + z * (zz * (((P[0] * zz) + P[1]) * zz + P[2]) / ((((zz + Q[1]) * zz + Q[2]) * zz + Q[3]) * zz + Q[4]))
    else:
        y = z
    if j & 2 == 2:
        y = -1 / y
    return -y if sign else y


# ---- scene reader ---------------------------------------------------------------------------------------------------------
def f32(tok):
    return float(np.float32(tok))              # strconv.ParseFloat(s, 32), widened (mesh.go:176)


def clamp01(v):                                # colour.NewRGBFromFloats, colour.go:33-35
    return max(0.0, min(v, 1.0))


def read_mtl(path):
    lib, cur = {}, None
    for line in open(path):
        t = line.split()
        if not t or t[0][0] == "#":
            continue
        if t[0] == "newmtl":
            cur = lib.setdefault(t[1], {"Ka": (0.0,) * 3, "Kd": (0.0,) * 3, "Ks": (0.0,) * 3, "Ns": 0.0})
        elif cur is not None and t[0] in ("Ka", "Kd", "Ks"):
            cur[t[0]] = tuple(f32(v) for v in t[1:4])
        elif cur is not None and t[0] == "Ns":
            cur["Ns"] = f32(t[1])
    return lib


def read_obj_scene(scene_json):
    """-> the `sc` dict pyref.Scene takes.  Vertices / normals are stored per face corner (no dedup: it does not change a
    single rendered bit, mesh.go:182-203 only shares equal values); face index = emission order."""
    root = os.path.dirname(scene_json)
    env = json.load(open(scene_json))
    meshes, mesh_of, obj_pos, obj_mesh = [], {}, [], []
    for o in env["objs"]:
        if o["model"] not in mesh_of:
            v, vn, faces, mats, mat_of = [], [], [], [], {}
            verts, norms = [], []
            lib, usemtl = {}, ""
            for line in open(os.path.join(root, o["model"])):
                t = line.split()
                if not t or t[0][0] == "#":
                    continue
                if t[0] == "mtllib":
                    lib = read_mtl(os.path.join(root, t[1]))
                elif t[0] == "usemtl":
                    usemtl = t[1] if len(t) > 1 else ""
                elif t[0] == "v":
                    v.append(tuple(f32(c) for c in t[1:4]))
                elif t[0] == "vn":
                    vn.append(tuple(f32(c) for c in t[1:4]))
                elif t[0] == "f":
                    if usemtl in lib:                                            # mesh.go:152-154
                        g = lib[usemtl]
                        mat = tuple(clamp01(c) for c in g["Ka"]) + tuple(clamp01(c) for c in g["Kd"]) + tuple(clamp01(c) for c in g["Ks"]) + (g["Ns"],)
                    else:                                                        # mesh.go:151
                        mat = (0x10 / 255.0,) * 3 + (1.0,) * 3 + (0.0,) * 3 + (0.0,)
                    if mat not in mat_of:
                        mat_of[mat] = len(mats)
                        mats.append(mat)
                    corners = []
                    for tok in t[1:]:
                        p = tok.split("/")
                        vi = int(p[0])
                        ni = int(p[2]) if len(p) > 2 and p[2] else None
                        corners.append((v[vi - 1 if vi > 0 else len(v) + vi], None if ni is None else vn[ni - 1 if ni > 0 else len(vn) + ni]))
                    assert len(corners) in (3, 4)
                    for tri in ([(0, 1, 2)] if len(corners) == 3 else [(0, 1, 2), (2, 3, 0)]):        # gwob's quad split
                        idx = []
                        for c in tri:
                            idx.append(len(verts))
                            verts.append(corners[c][0])
                            if corners[c][1] is not None:
                                norms.append(pyref.norm(corners[c][1]))                              # mesh.go:202
                        faces.append((idx[0], idx[1], idx[2], idx[0], idx[1], idx[2], mat_of[mat]))
            assert len(norms) in (0, len(verts))
            mesh_of[o["model"]] = len(meshes)
            meshes.append({"verts": verts, "normals": norms, "faces": faces, "materials": mats})
        obj_pos.append((o["pos"]["x"], o["pos"]["y"], o["pos"]["z"]))
        obj_mesh.append(mesh_of[o["model"]])
    lights = [(l["pos"]["x"], l["pos"]["y"], l["pos"]["z"], l["col"]["r"] / 255.0, l["col"]["g"] / 255.0, l["col"]["b"] / 255.0)
              for l in env["lights"]]                                           # colour.NewRGB, colour.go:28-30
    c = env["cam"]
    pos, d = (c["pos"]["x"], c["pos"]["y"], c["pos"]["z"]), (c["dir"]["x"], c["dir"]["y"], c["dir"]["z"])
    fwd = pyref.norm(d)                                                         # camera.go:39-41
    left = pyref.norm(pyref.cross(d, (0.0, 1.0, 0.0)))
    up = pyref.cross(left, fwd)
    return {"meshes": meshes, "obj_pos": obj_pos, "obj_mesh": obj_mesh, "lights": lights,
            "camera": list(pos) + list(fwd) + list(left) + list(up) + [c["fov"]]}


# ---- triangle.go:37-77 for one ray against every face of a mesh (numpy; same operation order as pyref.tri_intersection) ---
class FastScene(pyref.Scene):
    def __init__(self, sc):
        super().__init__(sc)
        self.np_meshes = []
        for m in self.meshes:
            v = np.array(m["v"], dtype=np.float64)
            f = np.array([t[0:3] for t in m["f"]], dtype=np.int64)
            P1, P2, P3 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
            self.np_meshes.append((P1, P2 - P1, P3 - P1))

    def candidates(self, mi, o, d):
        P1, e1, e2 = self.np_meshes[mi]
        nd = (-1 * d[0], -1 * d[1], -1 * d[2])

        def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
        def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
        E1, E2 = (e1[:, 0], e1[:, 1], e1[:, 2]), (e2[:, 0], e2[:, 1], e2[:, 2])
        with np.errstate(all="ignore"):
            c = cross(E2, nd)
            inc = dot(E1, c)
            por = (o[0] - P1[:, 0], o[1] - P1[:, 1], o[2] - P1[:, 2])
            r2 = dot(por, c) / inc
            r3 = dot(E1, cross(por, nd)) / inc
            r1 = 1.0 - r2 - r3
            s = dot(E1, cross(E2, por)) / inc
            ok = (inc != 0.0) & (r2 >= 0.0) & (r2 <= 1.0) & (r2 + r3 >= 0.0) & (r2 + r3 <= 1.0) & (r1 >= 0.0) & (r3 >= 0.0) & (s >= 0.0)
        return np.flatnonzero(ok)

    # object.go:63-110 restricted to the faces the numpy pass accepted (every other face returns None in the scalar code too)
    def object_intersection(self, pos, mesh, o, d):
        mi = next(k for k, m in enumerate(self.meshes) if m is mesh)
        ol = pyref.sub(o, pos)
        best = None
        for fi in self.candidates(mi, ol, d):
            f = mesh["f"][fi]
            P1, P2, P3 = mesh["v"][f[0]], mesh["v"][f[1]], mesh["v"][f[2]]
            hit = pyref.tri_intersection(P1, P2, P3, ol, d)
            assert hit is not None, "numpy and scalar Moller-Trumbore disagree"
            ip, (r1, r2, r3) = hit
            if mesh["n"]:
                N1, N2, N3 = mesh["n"][f[3]], mesh["n"][f[4]], mesh["n"][f[5]]
                normal = pyref.norm(pyref.add(pyref.add(pyref.scale(N1, r1), pyref.scale(N2, r2)), pyref.scale(N3, r3)))
            else:
                normal = pyref.norm(pyref.cross(pyref.sub(P2, P1), pyref.sub(P3, P1)))
            dist = pyref.length(pyref.sub(ol, ip))
            if best is None or dist < best[0]:
                best = (dist, ip, normal, mesh["m"][f[6]], int(fi))
        if best is None:
            return None
        return pyref.add(best[1], pos), best[2], best[3], best[4]


def render(S, W, H, tan_half, pixels):
    face = np.full(len(pixels), -1, dtype=np.int32)
    lit = np.zeros(len(pixels), dtype=np.uint32)
    col = np.zeros((len(pixels), 3), dtype=np.float64)
    for k, (i, j) in enumerate(pixels):
        hit, c, fi, _, l = S.trace_pixel(i, j, W, H, tan_half)
        if hit:
            face[k], lit[k], col[k] = fi, l, c
    rgb = np.array([pyref.quantise_rgb(c) for c in col], dtype=np.uint8)
    sdl = np.array([pyref.quantise_rgba16_shr8(c) for c in col], dtype=np.uint8)
    return face, lit, col, rgb, sdl


if __name__ == "__main__":
    sc = read_obj_scene(os.path.join(HERE, "example", "scene.json"))
    assert len(sc["meshes"][0]["faces"]) == 968
    tan_half = go_tan(sc["camera"][12] / 2.0)
    assert tan_half.hex() == float.fromhex("0x1.279a745228f6ap-1").hex()        # 0x3fe279a745228f6a
    S, slow = FastScene(sc), pyref.Scene(sc)
    rnd = random.Random(7)
    for _ in range(300):                                                        # the numpy candidate search == the scalar code
        i, j = rnd.randrange(100, 220), rnd.randrange(60, 180)
        assert S.trace_pixel(i, j, 320, 240, tan_half) == slow.trace_pixel(i, j, 320, 240, tan_half)
    # C1: every pixel, BulkTrace order (index i*H + j)
    W, H = 320, 240
    face, lit, col, rgb, sdl = render(S, W, H, tan_half, [(i, j) for i in range(W) for j in range(H)])
    hit = face >= 0
    np.savez_compressed(os.path.join(HERE, "c1_320x240_pyref.npz"), face=face.astype(np.int16), hit_index=np.flatnonzero(hit).astype(np.int32),
                        colour_bits=col[hit].view(np.uint64), rgb=rgb[hit], sdl=sdl[hit], lit=lit[hit].astype(np.uint8))
    print("C1: %d hit pixels" % hit.sum())
    # C2: lights[0] only, 1920x1080, every 97th pixel of the BulkTrace order
    sc2 = dict(sc)
    sc2["lights"] = sc["lights"][:1]
    S2 = FastScene(sc2)
    W, H = 1920, 1080
    idx = np.arange(5, W * H, 97)
    face, lit, col, rgb, sdl = render(S2, W, H, tan_half, [(int(k) // H, int(k) % H) for k in idx])
    np.savez_compressed(os.path.join(HERE, "c2_1080p_sample_pyref.npz"), index=idx.astype(np.int32), face=face.astype(np.int16),
                        colour_bits=col.view(np.uint64), rgb=rgb, sdl=sdl, lit=lit.astype(np.uint8))
    print("C2 sample: %d pixels, %d hits" % (len(idx), (face >= 0).sum()))
