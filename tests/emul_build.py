"""Builds and binds tests/emul/emul.cpp: a TEST-ONLY CPU replay of the device functions the sm_100a
kernels are made of (see the header of emul.cpp).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib(defs=()):
    """defs: extra -D options (e.g. ("-DRT_NODEQ=1",)): a second build of the same replay with a compile-time experiment switched on"""
    defs = tuple(defs)
    _SO = os.path.join(_HERE, "emul", "libemul%s.so" % ("".join("_" + d.lstrip("-D").replace("=", "") for d in defs)))
    _LIB = _LIBS.get(defs)
    if _LIB is None:
        src = os.path.join(_HERE, "emul", "emul.cpp")
        csrc = os.path.join(_HERE, "..", "distributed-raytracer_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("rt_types.h", "rt_bvh.cuh", "rt_traverse.cuh", "rt_exact.cuh", "rt_cull.h")] + \
            [os.path.join(_HERE, "emul", "host_builders_experiment.hpp")]
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", *defs,
                                   "-o", _SO, src])
        L = C.CDLL(_SO)
        L.emul_scene_create.restype = C.c_void_p
        L.emul_scene_create.argtypes = [C.POINTER(_lib.SceneDesc), C.c_int]
        L.emul_scene_create2.restype = C.c_void_p
        L.emul_scene_create2.argtypes = [C.POINTER(_lib.SceneDesc), C.c_int, C.c_int]
        L.emul_scene_destroy.argtypes = [C.c_void_p]
        L.emul_bvh_check.restype = C.c_int
        L.emul_bvh_check.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 3
        L.emul_render.restype = C.c_int
        L.emul_render.argtypes = [C.c_void_p, C.POINTER(_lib.Camera), C.c_void_p] + [C.c_uint32] * 8 + [C.c_void_p] * 6
        L.emul_render_tiles.restype = C.c_int
        L.emul_render_tiles.argtypes = [C.c_void_p, C.POINTER(_lib.Camera), C.c_void_p] + [C.c_uint32] * 8 + [C.c_void_p]
        for fn in (L.emul_go_max, L.emul_go_min, L.emul_go_pow):
            fn.restype = C.c_double
            fn.argtypes = [C.c_double, C.c_double]
        L.emul_morton63.restype = C.c_uint64
        L.emul_morton63.argtypes = [C.c_double] * 3
        _LIB = _LIBS[defs] = L
    return _LIB


class EmulScene:
    def __init__(self, scene, precision, max_leaf=4, max_leaf_any=0, defs=()):
        desc, keep = rt.make_scene_desc(scene, 0, precision)
        self.defs = tuple(defs)
        self.h = lib(self.defs).emul_scene_create2(C.byref(desc), max_leaf, max_leaf_any)
        self.scene = scene

    def bvh_check(self, mesh=0):
        out = [np.zeros(1, dtype=np.uint32) for _ in range(3)]
        rc = lib(self.defs).emul_bvh_check(self.h, mesh, *[o.ctypes.data for o in out])
        return rc, int(out[0][0]), int(out[1][0]), int(out[2][0])

    def render(self, cam13, lights, W, H, x=0, y=0, w=None, h=None, flags=0, tan_half_fov=None):
        from oracle import oracle as orc  # only for go_tan of the frame constant (test infrastructure)
        w = W - x if w is None else w
        h = H - y if h is None else h
        cam = rt.make_camera(cam13, orc.go_tan(float(cam13[12]) / 2.0) if tan_half_fov is None else tan_half_fov)
        lights = np.ascontiguousarray(lights, dtype=np.float64).reshape(-1, 6)
        n = w * h
        rgb = np.empty((n, 3), dtype=np.uint8)
        col = np.empty((n, 3), dtype=np.float64)
        face = np.empty(n, dtype=np.int32)
        obj = np.empty(n, dtype=np.int32)
        lit = np.empty(n, dtype=np.uint32)
        cnt = np.zeros(3, dtype=np.uint64)
        lib(self.defs).emul_render(self.h, C.byref(cam), lights.ctypes.data, lights.shape[0], W, H, x, y, w, h, flags,
                          rgb.ctypes.data, col.ctypes.data, face.ctypes.data, obj.ctypes.data, lit.ctypes.data, cnt.ctypes.data)
        return rgb, col, face, obj, lit, cnt

    def render_tiles(self, cam13, lights, W, H, tw, th, rank, n_ranks, out, flags=0):
        from oracle import oracle as orc
        cam = rt.make_camera(cam13, orc.go_tan(float(cam13[12]) / 2.0))
        lights = np.ascontiguousarray(lights, dtype=np.float64).reshape(-1, 6)
        lib(self.defs).emul_render_tiles(self.h, C.byref(cam), lights.ctypes.data, lights.shape[0], W, H, tw, th, rank, n_ranks, flags,
                                out.ctypes.data)
        return out

    def close(self):
        if self.h:
            lib(self.defs).emul_scene_destroy(self.h)
            self.h = None
