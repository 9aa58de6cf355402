"""Python handle over the C ABI (include/rt_b200.h): scene create / frame set / render region.

Mirrors the lifecycle of the reference worker (paths under /root/reference/):
  register()  worker/distributed/main.go:101-129  -> GpuScene(...)          (rt_scene_create)
  BulkTrace   worker/distributed/main.go:46-91    -> GpuScene.bulk_trace()  (rt_bulk_trace)
  draw()      worker/sequential/main.go:15-32     -> GpuScene.render_region()
Inputs are flat numpy arrays (the SoA flattening of state.Environment).
"""
import ctypes as C

import numpy as np

from . import _lib


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def make_scene_desc(scene, device=0, precision=_lib.FP32_FAST, flags=0):
    """scene: dict(meshes=[dict(verts, normals, faces, materials)], obj_pos, obj_mesh).
    Returns (SceneDesc, keepalive list)."""
    keep = []
    n = len(scene["meshes"])
    descs = (_lib.MeshDesc * max(1, n))()
    for k, m in enumerate(scene["meshes"]):
        v = _c(m["verts"], np.float64).reshape(-1, 3)
        nr = _c(m["normals"], np.float64).reshape(-1, 3)
        f = _c(m["faces"], np.uint32).reshape(-1, 7)
        mt = _c(m["materials"], np.float64).reshape(-1, 10)
        keep += [v, nr, f, mt]
        descs[k] = _lib.MeshDesc(v.ctypes.data, v.shape[0], nr.ctypes.data if nr.size else None, nr.shape[0],
                                 f.ctypes.data if f.size else None, f.shape[0], mt.ctypes.data if mt.size else None,
                                 mt.shape[0])
    op = _c(scene["obj_pos"], np.float64).reshape(-1, 3)
    om = _c(scene["obj_mesh"], np.uint32).reshape(-1)
    keep += [op, om, descs]
    d = _lib.SceneDesc(descs, n, op.ctypes.data if op.size else None, om.ctypes.data if om.size else None, op.shape[0],
                       device, precision, flags)
    return d, keep


def make_camera(cam13, tan_half_fov=float("nan")):
    cam13 = np.asarray(cam13, dtype=np.float64).reshape(13)
    c = _lib.Camera()
    for k in range(3):
        c.pos[k], c.forward[k], c.left[k], c.up[k] = cam13[k], cam13[3 + k], cam13[6 + k], cam13[9 + k]
    c.fov = cam13[12]
    c.tan_half_fov = tan_half_fov
    return c


class FrameState:
    """Per-frame mutable state (camera, lights, object positions) marshalled ONCE into the C structs of the ABI, so that the
    per-frame call is a single pointer hand-off (matters when a frame is ~0.1 ms)."""

    def __init__(self, cam13, lights, obj_pos=None, tan_half_fov=float("nan")):
        self.cam = make_camera(cam13, tan_half_fov)
        self._lights = _c(lights, np.float64).reshape(-1, 6)
        self.n_lights = self._lights.shape[0]
        self.lights_ptr = self._lights.ctypes.data if self._lights.size else None
        self._obj = None if obj_pos is None else _c(obj_pos, np.float64)
        self.obj_ptr = self._obj.ctypes.data if self._obj is not None else None

    def check(self, n_objects):
        """The ABI reads 3 * n_objects doubles behind obj_pos (it carries no length): refuse anything else here."""
        if self._obj is not None and self._obj.size != 3 * n_objects:
            raise ValueError("obj_pos holds %d values, the scene has %d objects (3 each)" % (self._obj.size, n_objects))


class GpuScene:
    def __init__(self, scene, device=0, precision=_lib.FP32_FAST, flags=0):
        self._L = _lib.load()
        desc, keep = make_scene_desc(scene, device, precision, flags)
        h = C.c_void_p()
        rc = self._L.rt_scene_create(C.byref(desc), C.byref(h))
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self.precision = precision
        self.n_objects = desc.n_objects
        self.n_meshes = desc.n_meshes
        if "camera" in scene and "lights" in scene:
            self.set_frame(scene["camera"], scene["lights"])

    @classmethod
    def from_handle(cls, handle, n_objects, device=0, precision=_lib.FP32_FAST):
        """Wraps an rt_scene* created elsewhere (host.gpu_scene_from_file: registration from OBJ text)."""
        self = cls.__new__(cls)
        self._L = _lib.load()
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.device, self.precision, self.n_objects, self.n_meshes = device, precision, n_objects, None
        return self

    def _check(self, rc):
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_last_error(self.h) or b"").decode())

    def _obj_pos(self, obj_pos):
        """obj_pos -> contiguous float64 array of exactly 3 * n_objects values (the ABI carries no length) or None."""
        if obj_pos is None:
            return None
        op = _c(obj_pos, np.float64)
        if op.size != 3 * self.n_objects:
            raise ValueError("obj_pos holds %d values, the scene has %d objects (3 each)" % (op.size, self.n_objects))
        return op

    def set_frame(self, cam13, lights, obj_pos=None, tan_half_fov=float("nan")):
        cam = make_camera(cam13, tan_half_fov)
        lights = _c(lights, np.float64).reshape(-1, 6)
        op = self._obj_pos(obj_pos)
        self._check(self._L.rt_frame_set(self.h, C.byref(cam), lights.ctypes.data if lights.size else None, lights.shape[0],
                                         op.ctypes.data if op is not None else None))

    def render_region(self, W, H, x=0, y=0, w=None, h=None, flags=0, out=None):
        w = W - x if w is None else w
        h = H - y if h is None else h
        if out is None:
            out = np.empty((w * h, 3), dtype=np.uint8)
        self._check(self._L.rt_render_region(self.h, W, H, x, y, w, h, out.ctypes.data, flags))
        return out

    def bulk_trace(self, cam13, lights, W, H, x, y, w, h, obj_pos=None, flags=0, out=None, tan_half_fov=float("nan")):
        cam = make_camera(cam13, tan_half_fov)
        lights = _c(lights, np.float64).reshape(-1, 6)
        op = self._obj_pos(obj_pos)
        if out is None:
            out = np.empty((w * h, 3), dtype=np.uint8)
        self._check(self._L.rt_bulk_trace(self.h, C.byref(cam), lights.ctypes.data if lights.size else None, lights.shape[0],
                                          op.ctypes.data if op is not None else None, W, H, x, y, w, h, out.ctypes.data, flags))
        return out

    def bulk_trace_encoded(self, cam13, lights, W, H, x, y, w, h, obj_pos=None, flags=0, out=None, tan_half_fov=float("nan")):
        """One BulkTrace down to the protobuf bytes of comms.TraceResults (rt_bulk_trace_encoded).  Returns a uint8 view of the
        message inside `out` (allocated at the worst-case size if not given)."""
        cam = make_camera(cam13, tan_half_fov)
        lights = _c(lights, np.float64).reshape(-1, 6)
        op = self._obj_pos(obj_pos)
        if out is None:
            out = np.empty(int(self._L.rt_trace_results_max_bytes(w * h)), dtype=np.uint8)
        n = C.c_uint64(0)
        self._check(self._L.rt_bulk_trace_encoded(self.h, C.byref(cam), lights.ctypes.data if lights.size else None, lights.shape[0],
                                                  op.ctypes.data if op is not None else None, W, H, x, y, w, h, out.ctypes.data,
                                                  out.nbytes, C.byref(n), flags))
        return out.reshape(-1)[:n.value]

    def encode_trace_results_device(self, d_rgb_ptr, n_pixels, d_out_ptr, capacity, d_len_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_encode_trace_results_device(self.h, d_rgb_ptr, n_pixels, d_out_ptr, capacity, d_len_ptr, stream_ptr, flags))

    def render_region_debug(self, W, H, x=0, y=0, w=None, h=None, flags=0):
        w = W - x if w is None else w
        h = H - y if h is None else h
        n = w * h
        rgb = np.empty((n, 3), dtype=np.uint8)
        col = np.empty((n, 3), dtype=np.float64)
        face = np.empty(n, dtype=np.int32)
        obj = np.empty(n, dtype=np.int32)
        lit = np.empty(n, dtype=np.uint32)
        self._check(self._L.rt_render_region_debug(self.h, W, H, x, y, w, h, rgb.ctypes.data, col.ctypes.data, face.ctypes.data,
                                                   obj.ctypes.data, lit.ctypes.data, flags))
        return rgb, col, face, obj, lit

    def render_region_device(self, W, H, x, y, w, h, d_out_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_render_region_device(self.h, W, H, x, y, w, h, d_out_ptr, stream_ptr, flags))

    def render_tiles_device(self, W, H, tile_w, tile_h, rank, n_ranks, d_out_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_render_tiles_device(self.h, W, H, tile_w, tile_h, rank, n_ranks, d_out_ptr, stream_ptr, flags))

    def render_tiles_peer(self, W, H, tile_w, tile_h, rank, n_ranks, d_framebuffer_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_render_tiles_peer(self.h, W, H, tile_w, tile_h, rank, n_ranks, d_framebuffer_ptr, stream_ptr, flags))

    def push_tiles_device(self, W, H, tile_w, tile_h, rank, n_ranks, d_tiles_ptr, d_framebuffer_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_push_tiles_device(self.h, W, H, tile_w, tile_h, rank, n_ranks, d_tiles_ptr, d_framebuffer_ptr, stream_ptr, flags))

    def render_push_tiles(self, frame, W, H, tile_w, tile_h, rank, n_ranks, d_tiles_ptr, d_framebuffer_ptr, stream_ptr, flags=0,
                          sync_slot_ptr=None, seq=0):
        """frame = FrameState (prebuilt ctypes camera + lights): one C call per frame.  sync_slot_ptr/seq: device-side frame
        hand-off through the slot's record in the gathering rank's memory (include/rt_b200.h)."""
        frame.check(self.n_objects)
        self._check(self._L.rt_render_push_tiles(self.h, C.byref(frame.cam), frame.lights_ptr, frame.n_lights, frame.obj_ptr, W, H, tile_w,
                                                 tile_h, rank, n_ranks, d_tiles_ptr, d_framebuffer_ptr, sync_slot_ptr, seq, stream_ptr, flags))

    def render_stripes_host(self, frame, W, H, stripe, rank, n_ranks, d_local_fb_ptr, h_framebuffer_ptr, stream_ptr, flags=0,
                            sync_slot_ptr=None, seq=0):
        """The multi-GPU end-to-end call of a rank: render its stripes, one strided D2H transfer into the host framebuffer."""
        frame.check(self.n_objects)
        self._check(self._L.rt_render_stripes_host(self.h, C.byref(frame.cam), frame.lights_ptr, frame.n_lights, frame.obj_ptr, W, H, stripe,
                                                   rank, n_ranks, d_local_fb_ptr, h_framebuffer_ptr, sync_slot_ptr, seq, stream_ptr, flags))

    def frame_complete_device(self, sync_slot_ptr, seq, n_ranks, release, stream_ptr):
        self._check(self._L.rt_frame_complete_device(self.h, sync_slot_ptr, seq, n_ranks, 1 if release else 0, stream_ptr))

    def frame_release_device(self, sync_slot_ptr, seq, stream_ptr):
        self._check(self._L.rt_frame_release_device(self.h, sync_slot_ptr, seq, stream_ptr))

    def set_frame_state(self, frame, lane=0):
        frame.check(self.n_objects)
        self._check(self._L.rt_frame_set_lane(self.h, lane, C.byref(frame.cam), frame.lights_ptr, frame.n_lights, frame.obj_ptr))

    def lane_count(self):
        return self._L.rt_lane_count(self.h)

    def untile_device(self, W, H, tile_w, tile_h, n_ranks, d_all_tiles_ptr, d_out_ptr, stream_ptr=None, flags=0):
        self._check(self._L.rt_untile_device(self.h, W, H, tile_w, tile_h, n_ranks, d_all_tiles_ptr, d_out_ptr, stream_ptr, flags))

    def stats(self):
        st = _lib.Stats()
        self._check(self._L.rt_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def bvh_info(self, mesh=0):
        bi = _lib.BvhInfo()
        self._check(self._L.rt_get_bvh_info(self.h, mesh, C.byref(bi)))
        return bi.as_dict()

    def synchronize(self):
        self._check(self._L.rt_synchronize(self.h))

    def cancel(self):
        self._check(self._L.rt_cancel(self.h))

    def close(self):
        if getattr(self, "h", None):
            self._L.rt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuGroup:
    """One worker process, several GPUs (rt_group_*): the scene is registered on every device and ONE bulk_trace call is
    answered by all of them (stripes of the region dealt round-robin, every GPU copies its stripes into the reply over its own
    PCIe link).  What master/pool sees is still one worker answering one BulkTrace (pool.go:148-197); no torch.distributed,
    no second process.  devices=None: every GPU of the box."""

    def __init__(self, scene, devices=None, precision=_lib.FP32_FAST):
        self._L = _lib.load()
        desc, keep = make_scene_desc(scene, 0, precision)
        h = C.c_void_p()
        devs = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._L.rt_group_create(C.byref(desc), devs.ctypes.data if devs is not None else None, 0 if devs is None else devs.size, C.byref(h))
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_group_last_error(None) or b"").decode())
        self.h = h
        self.n_objects = desc.n_objects
        self.n_devices = self._L.rt_group_device_count(self.h)

    def bulk_trace(self, cam13, lights, W, H, x, y, w, h, obj_pos=None, flags=0, out=None, tan_half_fov=float("nan")):
        cam = make_camera(cam13, tan_half_fov)
        lights = _c(lights, np.float64).reshape(-1, 6)
        op = None if obj_pos is None else _c(obj_pos, np.float64)
        if op is not None and op.size != 3 * self.n_objects:
            raise ValueError("obj_pos holds %d values, the scene has %d objects (3 each)" % (op.size, self.n_objects))
        if out is None:
            out = np.empty((w * h, 3), dtype=np.uint8)
        rc = self._L.rt_group_bulk_trace(self.h, C.byref(cam), lights.ctypes.data if lights.size else None, lights.shape[0],
                                         op.ctypes.data if op is not None else None, W, H, x, y, w, h, out.ctypes.data, flags)
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_group_last_error(self.h) or b"").decode())
        return out

    def close(self):
        if getattr(self, "h", None):
            self._L.rt_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tiles_per_rank(W, H, tile_w, tile_h, n_ranks):
    n_tiles = ((W + tile_w - 1) // tile_w) * ((H + tile_h - 1) // tile_h)
    return (n_tiles + n_ranks - 1) // n_ranks


class DeviceBuffer:
    """A library-allocated device buffer that can be shared with the other per-GPU processes of the box (CUDA IPC).
    Exposes __cuda_array_interface__ so torch.as_tensor(buf, device=...) gives a zero-copy uint8 view."""

    def __init__(self, device, nbytes=None, ipc_handle=None):
        self._L = _lib.load()
        self.device, self.nbytes, self.owner = device, nbytes, ipc_handle is None
        p = C.c_void_p()
        if ipc_handle is None:
            rc = self._L.rt_device_alloc(device, nbytes, C.byref(p))
        else:
            rc = self._L.rt_ipc_open(device, ipc_handle, C.byref(p))
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_last_error(None) or b"").decode())
        self.ptr = p.value

    def export(self):
        h = C.create_string_buffer(64)
        rc = self._L.rt_ipc_export(self.ptr, h)
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_last_error(None) or b"").decode())
        return h.raw

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3, "strides": None}

    def close(self):
        if getattr(self, "ptr", None):
            (self._L.rt_device_free if self.owner else self._L.rt_ipc_close)(self.device, self.ptr)
            self.ptr = None


class HostSharedBuffer:
    """A POSIX shared-memory segment (mapped by every per-GPU process of the box) pinned for this process's GPU: `.ptr` is the
    address the GPU's kernels use, `.array` the uint8 numpy view the host reads.  rank 0 creates (and finally unlinks) it, the
    others attach by name."""

    def __init__(self, device, nbytes=None, name=None):
        from multiprocessing import shared_memory
        self._L = _lib.load()
        self.device, self.owner = device, name is None
        if self.owner:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            try:                                   # the creator unlinks; keep Python's resource tracker from doing it at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.nbytes = nbytes if nbytes is not None else self.shm.size
        self.array = np.ndarray((self.nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        if self.owner:
            self.array[:] = 0
        self._host_ptr = self.array.ctypes.data
        p = C.c_void_p()
        rc = self._L.rt_host_register(device, self._host_ptr, self.shm.size, C.byref(p))
        if rc != 0:
            raise _lib.RtError(rc, (self._L.rt_last_error(None) or b"").decode())
        self.ptr = p.value

    @property
    def name(self):
        return self.shm.name

    def close(self):
        if getattr(self, "ptr", None):
            self._L.rt_host_unregister(self.device, self._host_ptr)
            self.ptr = None
            self.array = None
            try:
                self.shm.close()
            except BufferError:
                pass
            if self.owner:
                try:
                    self.shm.unlink()
                except FileNotFoundError:
                    pass
