"""ctypes wrapper over oracle/librt_oracle.so (CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT).

PARITY UNPINNED: see rt_oracle.h.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BRUTE, CULLED, RTREE = 0, 1, 2
Q_RGB255, Q_RGBA16_SHR8 = 0, 1
L_COLMAJOR, L_ROWMAJOR = 0, 1


class MeshDesc(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("n_verts", C.c_uint32),
                ("normals", C.c_void_p), ("n_normals", C.c_uint32),
                ("faces", C.c_void_p), ("n_faces", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("primary_rays", "shadow_rays", "hit_pixels", "tri_tests", "box_tests", "nearest_ties")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build():
    so = os.path.join(_HERE, "librt_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("rt_oracle.c", "rt_oracle.h", "Makefile")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_go_tan.restype = C.c_double
        L.orc_go_tan.argtypes = [C.c_double]
        L.orc_go_pow.restype = C.c_double
        L.orc_go_pow.argtypes = [C.c_double, C.c_double]
        for fn in (L.orc_go_max, L.orc_go_min):
            fn.restype = C.c_double
            fn.argtypes = [C.c_double, C.c_double]
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_build_rtrees.argtypes = [C.c_void_p]
        L.orc_frame_set.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_trace.restype = C.c_int
        L.orc_trace.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.orc_trace_pixel.restype = C.c_int
        L.orc_trace_pixel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_render_region.restype = C.c_int
        L.orc_render_region.argtypes = [C.c_void_p] + [C.c_int] * 12 + [C.c_void_p] * 4
        L.orc_triangle_intersection.restype = C.c_int
        L.orc_triangle_intersection.argtypes = [C.c_void_p] * 6
        L.orc_box_intersect.restype = C.c_int
        L.orc_box_intersect.argtypes = [C.c_void_p] * 4
        L.orc_new_camera.restype = C.c_int
        L.orc_new_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_primary_ray.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2
        L.orc_encode_trace_results.restype = C.c_uint64
        L.orc_encode_trace_results.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def go_tan(x):
    return lib().orc_go_tan(float(x))


def go_pow(x, y):
    return lib().orc_go_pow(float(x), float(y))


def encode_trace_results(rgb):
    """comms.TraceResults wire bytes of an (n,3) uint8 array in BulkTrace order (rt_oracle.h: orc_encode_trace_results)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    n = rgb.shape[0]
    out = np.empty(11 * n, dtype=np.uint8)
    m = lib().orc_encode_trace_results(_p(rgb) if n else None, n, _p(out) if n else None)
    return out[:m].copy()


def triangle_intersection(tri, o, d):
    tri = np.ascontiguousarray(tri, dtype=np.float64).reshape(9)
    o = np.ascontiguousarray(o, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    ip = np.zeros(3)
    bary = np.zeros(3)
    t = np.zeros(1)
    hit = lib().orc_triangle_intersection(_p(tri), _p(o), _p(d), _p(ip), _p(bary), _p(t))
    return bool(hit), ip, bary, float(t[0])


def box_intersect(bmin, bmax, o, d):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (bmin, bmax, o, d)]
    return bool(lib().orc_box_intersect(*[_p(v) for v in a]))


def primary_ray(i, j, W, H, cam13):
    cam13 = np.ascontiguousarray(cam13, dtype=np.float64)
    d = np.zeros(3)
    lib().orc_primary_ray(i, j, W, H, _p(cam13), _p(d))
    return d


class Scene:
    """scene dict layout: see oracle/scene_loader.environment_from_file."""

    def __init__(self, scene):
        L = lib()
        self._keep = []
        descs = (MeshDesc * max(1, len(scene["meshes"])))()
        for k, m in enumerate(scene["meshes"]):
            v = np.ascontiguousarray(m["verts"], dtype=np.float64)
            n = np.ascontiguousarray(m["normals"], dtype=np.float64)
            f = np.ascontiguousarray(m["faces"], dtype=np.uint32)
            mt = np.ascontiguousarray(m["materials"], dtype=np.float64)
            self._keep += [v, n, f, mt]
            descs[k] = MeshDesc(v.ctypes.data, v.shape[0], n.ctypes.data if n.size else None, n.shape[0] if n.size else 0,
                                f.ctypes.data, f.shape[0], mt.ctypes.data, mt.shape[0])
        op = np.ascontiguousarray(scene["obj_pos"], dtype=np.float64)
        om = np.ascontiguousarray(scene["obj_mesh"], dtype=np.uint32)
        self.n_obj = op.shape[0]
        self.h = L.orc_scene_create(C.cast(descs, C.c_void_p), len(scene["meshes"]), _p(op), _p(om), self.n_obj)
        self.set_frame(scene["camera"], scene["lights"])

    def set_frame(self, cam13, lights, obj_pos=None, tan_half_fov=float("nan")):
        cam13 = np.ascontiguousarray(cam13, dtype=np.float64)
        lights = np.ascontiguousarray(lights, dtype=np.float64).reshape(-1, 6)
        op = None if obj_pos is None else np.ascontiguousarray(obj_pos, dtype=np.float64)
        lib().orc_frame_set(self.h, _p(cam13), float(tan_half_fov), _p(lights), lights.shape[0], _p(op))

    def build_rtrees(self):
        lib().orc_build_rtrees(self.h)

    def rtree_path_check(self, mesh, face, o, d):
        """(rectangles on the R-tree path to `face`, index of the first one box.go's predicate rejects or -1, that rectangle)"""
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        first = C.c_int(-1)
        bmin, bmax = np.zeros(3), np.zeros(3)
        lib().orc_rtree_path_check.restype = C.c_int
        lib().orc_rtree_path_check.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        n = lib().orc_rtree_path_check(self.h, mesh, face, _p(o), _p(d), C.byref(first), _p(bmin), _p(bmax))
        return n, first.value, bmin, bmax

    def trace(self, o, d, mode=CULLED):
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        ip, nrm = np.zeros(3), np.zeros(3)
        face, obj, mat = (np.zeros(1, dtype=np.int32) for _ in range(3))
        hit = lib().orc_trace(self.h, mode, _p(o), _p(d), _p(ip), _p(nrm), _p(face), _p(obj), _p(mat))
        return bool(hit), ip, nrm, int(face[0]), int(obj[0]), int(mat[0])

    def trace_pixel(self, i, j, W, H, mode=CULLED):
        col = np.zeros(3)
        face, obj = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        lit = np.zeros(1, dtype=np.uint32)
        hit = lib().orc_trace_pixel(self.h, mode, i, j, W, H, _p(col), _p(face), _p(obj), _p(lit))
        return bool(hit), col, int(face[0]), int(obj[0]), int(lit[0])

    def render_region(self, W, H, x=0, y=0, w=None, h=None, mode=CULLED, layout=L_COLMAJOR, quant=Q_RGB255,
                      n_threads=0, step=1, phase=0, want_colour=True, want_face=True):
        w = W if w is None else w
        h = H if h is None else h
        rgb = np.zeros((w * h, 3), dtype=np.uint8)
        col = np.zeros((w * h, 3), dtype=np.float64) if want_colour else None
        face = np.full(w * h, -1, dtype=np.int32) if want_face else None
        st = Stats()
        lib().orc_render_region(self.h, mode, W, H, x, y, w, h, layout, quant, n_threads, step, phase,
                                _p(rgb), _p(col), _p(face), C.cast(C.pointer(st), C.c_void_p))
        return rgb, col, face, st.as_dict()

    def close(self):
        if self.h:
            lib().orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
