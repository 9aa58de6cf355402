"""Independent Python restatement of the reference's scene loading (TEST INFRASTRUCTURE).

PARITY UNPINNED: gwob (github.com/mwindels/gwob, a fork of udhos/gwob; no version pinned, source
absent from /root/reference) does the OBJ/MTL parsing in the reference.  Its published behaviour is
restated here: coordinates parsed with 32-bit precision then widened (mesh.go:176 `Coord64`),
faces of 3 or 4 vertices, quad -> (v0 v1 v2) + (v2 v3 v0), one group per o/g/usemtl run, faces
emitted in file order.

Follows (paths under /root/reference/):
  shared/state/environment.go:162-234  EnvironmentFromFile
  shared/state/mesh.go:109-213         MeshFromFile (value-keyed dedup of vertices / normals / materials)
  shared/state/camera.go:35-44         NewCamera
  shared/colour/colour.go:28-35        NewRGB / NewRGBFromFloats

Used only by tests/ to cross-check the product's C++ loader and to feed the oracle.
"""
import json
import math
import os

import numpy as np


def _f32(tok):
    # strconv.ParseFloat(tok, 32) -> float32 -> float64
    return float(np.float32(tok))


def parse_mtl(path):
    lib = {}
    cur = None
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "newmtl":
                cur = {"Ka": [0.0, 0.0, 0.0], "Kd": [0.0, 0.0, 0.0], "Ks": [0.0, 0.0, 0.0], "Ns": 0.0}
                lib[t[1] if len(t) > 1 else ""] = cur
            elif cur is not None and t[0] in ("Ka", "Kd", "Ks"):
                cur[t[0]] = [_f32(t[1]), _f32(t[2]), _f32(t[3])]
            elif cur is not None and t[0] == "Ns":
                cur["Ns"] = _f32(t[1])
    return lib


def _clamp01(x):
    return max(0.0, min(float(x), 1.0))


def _material_from(lib, name):
    if name in lib:  # mesh.go:152-154
        m = lib[name]
        return tuple([_clamp01(c) for c in m["Ka"]] + [_clamp01(c) for c in m["Kd"]] +
                     [_clamp01(c) for c in m["Ks"]] + [float(m["Ns"])])
    # mesh.go:151 default material: Ka = NewRGB(0x10,...), Kd = NewRGB(0xFF,...), Ks = 0, Ns = 0
    ka = 0x10 / 255.0
    return (ka, ka, ka, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0)


def parse_obj(path):
    """Returns (groups, mtllib); each group = (usemtl, [tri, ...]); tri = ((v,vn),(v,vn),(v,vn)) with
    v / vn already resolved to float32-origin coordinate triples (vn None if absent)."""
    vs, vns = [], []
    groups = []
    cur = {"usemtl": "", "tris": []}
    groups.append(cur)
    mtllib = ""
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            k = t[0]
            if k == "v":
                vs.append((_f32(t[1]), _f32(t[2]), _f32(t[3])))
            elif k == "vn":
                vns.append((_f32(t[1]), _f32(t[2]), _f32(t[3])))
            elif k == "mtllib":
                mtllib = t[1]
            elif k in ("o", "g"):
                if cur["tris"]:
                    cur = {"usemtl": cur["usemtl"], "tris": []}
                    groups.append(cur)
            elif k == "usemtl":
                name = t[1] if len(t) > 1 else ""
                if cur["tris"]:
                    cur = {"usemtl": name, "tris": []}
                    groups.append(cur)
                else:
                    cur["usemtl"] = name
            elif k == "f":
                corners = []
                for tok in t[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0])
                    vi = vi - 1 if vi > 0 else len(vs) + vi
                    ni = None
                    if len(parts) >= 3 and parts[2] != "":
                        ni = int(parts[2])
                        ni = ni - 1 if ni > 0 else len(vns) + ni
                    corners.append((vs[vi], vns[ni] if ni is not None else None))
                if len(corners) == 3:
                    cur["tris"].append((corners[0], corners[1], corners[2]))
                elif len(corners) == 4:
                    cur["tris"].append((corners[0], corners[1], corners[2]))
                    cur["tris"].append((corners[2], corners[3], corners[0]))
                else:
                    raise ValueError("face with %d vertices unsupported (gwob: triangles and quads only)" % len(corners))
    return [g for g in groups if g["tris"]], mtllib


def _go_norm(v):
    # vector.go:50-53 ; Python floats are IEEE binary64 and never fused
    mag = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    if mag == 0.0:        # Go divides without trapping: 0/0 = NaN, x/0 = +-Inf (a zero `vn` line in the OBJ file)
        with np.errstate(all="ignore"):                      # the hardware's own NaN / Inf (DIVSD), as in the Go binary
            return tuple(float(np.float64(c) / np.float64(mag)) for c in v)
    return (v[0] / mag, v[1] / mag, v[2] / mag)


def mesh_from_file(path):
    """mesh.go:109-213 -> dict(verts, normals, faces, materials) numpy arrays."""
    groups, mtllib = parse_obj(path)
    lib = {}
    if mtllib:
        cand = os.path.join(os.path.dirname(path), mtllib.lstrip("/\\"))
        lib = parse_mtl(cand if os.path.exists(cand) else mtllib)
    has_normals = any(c[1] is not None for g in groups for tri in g["tris"] for c in tri)
    verts, vmap = [], {}
    norms, nmap = [], {}
    mats, mmap = [], {}
    faces = []
    for g in groups:
        mat = _material_from(lib, g["usemtl"])
        if mat not in mmap:
            mmap[mat] = len(mats)
            mats.append(mat)
        mi = mmap[mat]
        for tri in g["tris"]:
            fv, fn = [], []
            for (v, n) in tri:
                # Go map keys use ==, so -0.0 and +0.0 collide; Python dict/tuple hashing does the same.
                if v not in vmap:
                    vmap[v] = len(verts)
                    verts.append(v)
                fv.append(vmap[v])
                if has_normals:
                    if n is None:
                        n = (0.0, 0.0, 0.0)
                    if n not in nmap:  # keyed by the UN-normalised value (mesh.go:197)
                        nmap[n] = len(norms)
                        norms.append(_go_norm(n))  # stored normalised (mesh.go:202)
                    fn.append(nmap[n])
                else:
                    fn.append(0)
            faces.append(fv + fn + [mi])
    return {
        "verts": np.asarray(verts, dtype=np.float64).reshape(-1, 3),
        "normals": np.asarray(norms, dtype=np.float64).reshape(-1, 3),
        "faces": np.asarray(faces, dtype=np.uint32).reshape(-1, 7),
        "materials": np.asarray(mats, dtype=np.float64).reshape(-1, 10),
    }


def new_camera(pos, direction, fov):
    """camera.go:35-44 -> 13 doubles: pos, forward, left, up, fov."""
    d = [float(x) for x in direction]
    gup = (0.0, 1.0, 0.0)

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    c = cross(d, gup)
    if c[0] == 0.0 and c[1] == 0.0 and c[2] == 0.0:
        raise ValueError("Camera dir is parallel to global up")
    forward = _go_norm(d)
    left = _go_norm(c)
    up = cross(left, forward)
    return np.asarray(list(pos) + list(forward) + list(left) + list(up) + [float(fov)], dtype=np.float64)


def environment_from_file(path):
    """environment.go:162-234."""
    with open(path) as fh:
        js = json.load(fh)
    meshes, mesh_index = [], {}
    obj_pos, obj_mesh = [], []
    for o in js.get("objs", []):
        model = o["model"]
        if model not in mesh_index:
            cand = os.path.join(os.path.dirname(path), model.lstrip("/\\"))
            mesh_index[model] = len(meshes)
            meshes.append(mesh_from_file(cand if os.path.exists(cand) else model))
        obj_mesh.append(mesh_index[model])
        p = o["pos"]
        obj_pos.append([p["x"], p["y"], p["z"]])
    lights = []
    for l in js.get("lights", []):
        p, c = l["pos"], l["col"]
        lights.append([p["x"], p["y"], p["z"], c["r"] / 255.0, c["g"] / 255.0, c["b"] / 255.0])  # colour.go:28-30
    cam = js["cam"]
    cam13 = new_camera([cam["pos"][k] for k in "xyz"], [cam["dir"][k] for k in "xyz"], cam["fov"])
    return {
        "meshes": meshes,
        "obj_pos": np.asarray(obj_pos, dtype=np.float64).reshape(-1, 3),
        "obj_mesh": np.asarray(obj_mesh, dtype=np.uint32),
        "lights": np.asarray(lights, dtype=np.float64).reshape(-1, 6),
        "camera": cam13,
    }
