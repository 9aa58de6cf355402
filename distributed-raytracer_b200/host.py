"""ctypes binding of librt_host.so: the C++ mirror of the reference's Go host side for the render path
(host/state.hpp, host/gpu_worker.hpp).  Same names and argument meaning as the reference:

  EnvironmentFromFile   /root/reference/shared/state/environment.go:162-234
  Tracer.BulkTrace      /root/reference/worker/distributed/main.go:46-91
  draw                  /root/reference/worker/sequential/main.go:15-32
  partition             /root/reference/master/main.go:54-91
"""
import ctypes as C
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "librt_host.so")
_H = None


def load():
    global _H
    if _H is not None:
        return _H
    _lib.load()  # librt_b200.so first (librt_host.so links against it through $ORIGIN)
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError("%s not found: build it with `python __graft_entry__.py build`" % HOST_LIB_PATH)
    H = C.CDLL(HOST_LIB_PATH)
    vp, u32 = C.c_void_p, C.c_uint32
    H.rth_last_error.restype = C.c_char_p
    H.rth_environment_from_file.restype = vp
    H.rth_environment_from_file.argtypes = [C.c_char_p]
    H.rth_environment_free.argtypes = [vp]
    H.rth_environment_desc.restype = C.POINTER(_lib.SceneDesc)
    H.rth_environment_desc.argtypes = [vp]
    H.rth_environment_n_lights.restype = u32
    H.rth_environment_n_lights.argtypes = [vp]
    H.rth_environment_frame.argtypes = [vp, C.POINTER(_lib.Camera), vp, vp]
    H.rth_camera_move.argtypes = [vp, C.c_double] + [C.c_int] * 6
    H.rth_camera_yaw.argtypes = [vp, C.c_double]
    H.rth_camera_pitch.argtypes = [vp, C.c_double]
    H.rth_camera_set.argtypes = [vp, vp, vp, C.c_double]
    H.rth_tracer_new.restype = vp
    H.rth_tracer_new.argtypes = [vp, u32, u32, C.c_int, u32]
    H.rth_tracer_free.argtypes = [vp]
    H.rth_tracer_handle.restype = vp
    H.rth_tracer_handle.argtypes = [vp]
    H.rth_bulk_trace.restype = C.c_int
    H.rth_bulk_trace.argtypes = [vp, vp, u32, u32, u32, u32, vp, u32]
    H.rth_bulk_trace_encoded.restype = C.c_longlong
    H.rth_bulk_trace_encoded.argtypes = [vp, vp, u32, u32, u32, u32, vp, u32]
    H.rth_draw.restype = C.c_int
    H.rth_draw.argtypes = [vp, vp, vp, u32]
    H.rth_partition.restype = u32
    H.rth_partition.argtypes = [u32, u32, u32, vp, u32]
    _H = H
    return H


def _err():
    return (load().rth_last_error() or b"").decode()


class Environment:
    """state.Environment loaded by state.EnvironmentFromFile (environment.go:162-234)."""

    def __init__(self, path):
        H = load()
        self.h = H.rth_environment_from_file(os.fsencode(path))
        if not self.h:
            raise RuntimeError("EnvironmentFromFile: " + _err())

    def scene_arrays(self):
        """The SoA flattening (Environment::Flatten) as numpy copies, in the layout GpuScene takes."""
        d = load().rth_environment_desc(self.h).contents
        meshes = []
        for k in range(d.n_meshes):
            m = d.meshes[k]

            def arr(ptr, ctype, shape):
                if not ptr or shape[0] == 0:
                    return np.zeros(shape, dtype=np.dtype(ctype))
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape).copy()

            meshes.append({"verts": arr(m.verts, C.c_double, (m.n_verts, 3)),
                           "normals": arr(m.normals, C.c_double, (m.n_normals, 3)),
                           "faces": arr(m.faces, C.c_uint32, (m.n_faces, 7)),
                           "materials": arr(m.materials, C.c_double, (m.n_materials, 10))})
        n = d.n_objects
        op = np.ctypeslib.as_array(C.cast(d.obj_pos, C.POINTER(C.c_double)), (n, 3)).copy() if n else np.zeros((0, 3))
        om = np.ctypeslib.as_array(C.cast(d.obj_mesh, C.POINTER(C.c_uint32)), (n,)).copy() if n else np.zeros(0, np.uint32)
        cam13, tan_half, lights, _ = self.frame()
        return {"meshes": meshes, "obj_pos": op, "obj_mesh": om, "lights": lights, "camera": cam13, "tan_half_fov": tan_half}

    def frame(self):
        """(cam13, tan_half_fov, lights[n,6], obj_pos[n,3]) of the mutable state (environment.go:65-69)."""
        H = load()
        nl = H.rth_environment_n_lights(self.h)
        nobj = H.rth_environment_desc(self.h).contents.n_objects
        cam = _lib.Camera()
        lights = np.zeros((nl, 6), dtype=np.float64)
        op = np.zeros((nobj, 3), dtype=np.float64)
        H.rth_environment_frame(self.h, C.byref(cam), lights.ctypes.data, op.ctypes.data)
        cam13 = np.array(list(cam.pos) + list(cam.forward) + list(cam.left) + list(cam.up) + [cam.fov])
        return cam13, cam.tan_half_fov, lights, op

    # camera.go:62-154
    def cam_move(self, distance, forward=False, backward=False, leftward=False, rightward=False, upward=False, downward=False):
        load().rth_camera_move(self.h, distance, *[int(b) for b in (forward, backward, leftward, rightward, upward, downward)])

    def cam_yaw(self, theta):
        load().rth_camera_yaw(self.h, theta)

    def cam_pitch(self, theta):
        load().rth_camera_pitch(self.h, theta)

    def cam_set(self, pos, direction, fov):
        p = np.ascontiguousarray(pos, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        load().rth_camera_set(self.h, p.ctypes.data, d.ctypes.data, float(fov))

    def close(self):
        if getattr(self, "h", None):
            load().rth_environment_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tracer:
    """worker/distributed/main.go:27-32 Tracer, GPU-backed (register() == construction)."""

    def __init__(self, env, screen_width, screen_height, device=0, precision=_lib.FP32_FAST):
        self.env = env
        self.screen_width, self.screen_height = screen_width, screen_height
        self.t = load().rth_tracer_new(env.h, screen_width, screen_height, device, precision)
        if not self.t:
            raise RuntimeError("Tracer: " + _err())

    def BulkTrace(self, x, y, width, height, flags=0, out=None):
        """distributed/main.go:46-91: results[i*height + j] = (r, g, b)."""
        if out is None:
            out = np.empty((width * height, 3), dtype=np.uint8)
        if load().rth_bulk_trace(self.t, self.env.h, x, y, width, height, out.ctypes.data, flags) != 0:
            raise RuntimeError("BulkTrace: " + _err())
        return out

    def BulkTraceEncoded(self, x, y, width, height, flags=0):
        """The same order answered as the protobuf wire bytes of comms.TraceResults (comms.proto:34-41), encoded on the GPU."""
        out = np.empty(11 * width * height, dtype=np.uint8)
        n = load().rth_bulk_trace_encoded(self.t, self.env.h, x, y, width, height, out.ctypes.data, flags)
        if n < 0:
            raise RuntimeError("BulkTraceEncoded: " + _err())
        return out[:n]

    def draw(self, quantiser=_lib.QUANT_RGBA16_SHR8):
        """sequential/main.go:15-32: the whole screen, row-major [H, W, 3]."""
        surf = np.empty((self.screen_height, self.screen_width, 3), dtype=np.uint8)
        if load().rth_draw(self.t, self.env.h, surf.ctypes.data, quantiser) != 0:
            raise RuntimeError("draw: " + _err())
        return surf

    def handle(self):
        return load().rth_tracer_handle(self.t)

    def close(self):
        if getattr(self, "t", None):
            load().rth_tracer_free(self.t)
            self.t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gpu_scene_from_file(scene_path, device=0, precision=_lib.FP32_FAST):
    """register() straight from the scene file -- every model is handed to the GPU as the TEXT of its OBJ file and
    parsed there (rt_scene_create_obj); the host reads bytes, scans for mtllib / usemtl lines and parses the MTL file.
    Returns (GpuScene, cam13, tan_half_fov, lights[n,6], obj_pos[n,3]); the frame state is already set on the scene."""
    from .gpu_scene import GpuScene
    H_ = load()
    H_.rth_gpu_scene_from_file.restype = C.c_void_p
    H_.rth_gpu_scene_from_file.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    cam = _lib.Camera()
    lights = np.zeros((64, 6), dtype=np.float64)
    op = np.zeros((64, 3), dtype=np.float64)
    nl, nobj = C.c_uint32(0), C.c_uint32(0)
    h = H_.rth_gpu_scene_from_file(os.fsencode(scene_path), device, precision, C.byref(cam), lights.ctypes.data, 64, C.byref(nl),
                                   op.ctypes.data, 64, C.byref(nobj))
    if not h:
        raise RuntimeError("GpuSceneFromFile: " + _err())
    g = GpuScene.from_handle(h, nobj.value, device, precision)
    cam13 = np.array(list(cam.pos) + list(cam.forward) + list(cam.left) + list(cam.up) + [cam.fov])
    lights, op = lights[:nl.value].copy(), op[:nobj.value].copy()
    g.set_frame(cam13, lights, tan_half_fov=cam.tan_half_fov)
    return g, cam13, cam.tan_half_fov, lights, op


def obj_text_probe(obj_path):
    """The host half of the OBJ ingest alone (file read + mtllib / usemtl scan + MTL parse): (bytes, n_materials, n_usemtl)."""
    H_ = load()
    H_.rth_obj_text_probe.restype = C.c_longlong
    H_.rth_obj_text_probe.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    nm, nu = C.c_uint32(0), C.c_uint32(0)
    n = H_.rth_obj_text_probe(os.fsencode(obj_path), C.byref(nm), C.byref(nu))
    if n < 0:
        raise RuntimeError("ObjTextFromFile: " + _err())
    return int(n), nm.value, nu.value


def partition_balanced(width, height, workers, last_rgb, hit_weight=30.0):
    """partition() with the cuts placed by the cost of the PREVIOUS frame (`last_rgb`: width*height*3 bytes in
    BulkTrace order; a hit pixel costs 1 + hit_weight, a miss 1).  List of (x, y, w, h)."""
    H_ = load()
    rgb = np.ascontiguousarray(last_rgb, dtype=np.uint8).reshape(-1)
    if rgb.size != width * height * 3:
        raise ValueError("last_rgb must hold width*height*3 bytes")
    cap = max(4 * workers, 16)
    out = np.zeros((cap, 4), dtype=np.uint32)
    H_.rth_partition_balanced.restype = C.c_uint32
    H_.rth_partition_balanced.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_double, C.c_void_p, C.c_uint32]
    n = H_.rth_partition_balanced(width, height, workers, rgb.ctypes.data, float(hit_weight), out.ctypes.data, cap)
    return [tuple(int(v) for v in out[k]) for k in range(n)]


def partition(width, height, workers):
    """master/main.go:54-91: list of (x, y, w, h)."""
    cap = max(4 * workers, 16)
    out = np.zeros((cap, 4), dtype=np.uint32)
    n = load().rth_partition(width, height, workers, out.ctypes.data, cap)
    return [tuple(int(v) for v in out[k]) for k in range(n)]
