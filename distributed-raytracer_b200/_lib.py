"""ctypes binding of librt_b200.so (include/rt_b200.h).  Product code: there is NO fallback --
if the shared library (hand-written sm_100a kernels) is missing or does not load, importing the
GPU entry points raises; nothing here ever routes to a CPU implementation."""
import ctypes as C
import os

# Frames in flight live on separate CUDA streams; with the default 8 hardware queues several streams share a queue and a
# queue stalls at its first entry that has to wait (e.g. the consumer behind a frame hand-off wait).  Must be set before the
# CUDA context exists, hence at import time; an explicit setting of the user wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")   # override: kernel-variant experiments

FP64_EXACT, FP32_FAST = 0, 1
LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0x0, 0x1
QUANT_RGB255, QUANT_RGBA16_SHR8 = 0x0, 0x2
FLAG_COUNT, FLAG_BRUTE_FORCE, FLAG_SIMPLE, FLAG_PERSISTENT = 0x4, 0x8, 0x10, 0x20
FLAG_FUSED, FLAG_PIPELINE, FLAG_NO_CULL, FLAG_SM_AFFINE = 0x40, 0x80, 0x100, 0x200


def LANE(k):
    """RT_LANE(k): lane selector of the device API, carried in bits 24..31 of `flags`."""
    return int(k) << 24

RT_OK, RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_OOM, RT_ERR_STATE, RT_ERR_CANCELLED = 0, -1, -2, -3, -4, -5

# every symbol include/rt_b200.h declares
SYNC_SLOT_BYTES = 256            # include/rt_b200.h: RT_SYNC_SLOT_BYTES

ABI_SYMBOLS = (
    "rt_scene_create", "rt_scene_create_obj", "rt_scene_destroy", "rt_last_error", "rt_abi_version", "rt_device_count", "rt_frame_set", "rt_frame_set_lane", "rt_lane_count",
    "rt_render_region", "rt_bulk_trace", "rt_render_region_debug", "rt_render_region_device", "rt_render_tiles_device",
    "rt_untile_device", "rt_render_tiles_peer", "rt_push_tiles_device", "rt_render_push_tiles", "rt_render_stripes_host", "rt_frame_complete_device", "rt_frame_release_device",
    "rt_trace_results_max_bytes", "rt_encode_trace_results_device", "rt_bulk_trace_encoded", "rt_device_alloc", "rt_device_free", "rt_ipc_export", "rt_ipc_open",
    "rt_ipc_close", "rt_measure_fma_peak", "rt_get_stats", "rt_get_bvh_info", "rt_synchronize", "rt_cancel", "rt_host_alloc", "rt_host_free", "rt_host_register", "rt_host_unregister",
    "rt_group_create", "rt_group_destroy", "rt_group_device_count", "rt_group_scene", "rt_group_last_error", "rt_group_bulk_trace",
)


class MeshDesc(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("n_verts", C.c_uint32),
                ("normals", C.c_void_p), ("n_normals", C.c_uint32),
                ("faces", C.c_void_p), ("n_faces", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32)]


class ObjMeshDesc(C.Structure):
    _fields_ = [("obj_text", C.c_void_p), ("obj_bytes", C.c_uint64), ("materials", C.c_void_p), ("n_materials", C.c_uint32),
                ("usemtl_offset", C.c_void_p), ("usemtl_material", C.c_void_p), ("n_usemtl", C.c_uint32), ("default_material", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(MeshDesc)), ("n_meshes", C.c_uint32),
                ("obj_pos", C.c_void_p), ("obj_mesh", C.c_void_p), ("n_objects", C.c_uint32),
                ("device", C.c_int32), ("precision", C.c_uint32), ("flags", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("forward", C.c_double * 3), ("left", C.c_double * 3), ("up", C.c_double * 3),
                ("fov", C.c_double), ("tan_half_fov", C.c_double)]


class Light(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("col", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("hit_pixels", C.c_uint64),
                ("nodes_fetched", C.c_uint64), ("tris_fetched", C.c_uint64), ("nodes_primary", C.c_uint64),
                ("tris_primary", C.c_uint64), ("nodes_shade", C.c_uint64), ("tris_shade", C.c_uint64), ("node_bytes", C.c_uint64),
                ("tri_bytes", C.c_uint64), ("ms_primary", C.c_float), ("ms_shade", C.c_float), ("ms_write", C.c_float),
                ("ms_total", C.c_float), ("occluded_fraction", C.c_float), ("any_defer", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class BvhInfo(C.Structure):
    _fields_ = [("n_faces", C.c_uint64), ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64), ("max_depth", C.c_uint32),
                ("max_leaf", C.c_uint32), ("ms_build", C.c_float), ("sah_cost", C.c_float),
                ("n_nodes_any", C.c_uint64), ("n_leaves_any", C.c_uint64), ("max_depth_any", C.c_uint32), ("max_leaf_any", C.c_uint32),
                ("ms_ingest", C.c_float), ("n_verts_ingested", C.c_uint32), ("n_normals_ingested", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rt_b200 error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def load():
    """Loads librt_b200.so; raises (loudly) if it is missing -- there is no CPU fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a); "
                          "this backend has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    u32, vp = C.c_uint32, C.c_void_p
    L.rt_scene_create.restype = C.c_int
    L.rt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp)]
    L.rt_scene_create_obj.restype = C.c_int
    L.rt_scene_create_obj.argtypes = [C.POINTER(ObjMeshDesc), u32, vp, vp, u32, C.c_int32, u32, C.POINTER(vp)]
    L.rt_scene_destroy.restype = None
    L.rt_scene_destroy.argtypes = [vp]
    L.rt_last_error.restype = C.c_char_p
    L.rt_last_error.argtypes = [vp]
    L.rt_abi_version.restype = C.c_int
    L.rt_device_count.restype = C.c_int
    L.rt_frame_set.restype = C.c_int
    L.rt_frame_set.argtypes = [vp, C.POINTER(Camera), vp, u32, vp]
    L.rt_frame_set_lane.restype = C.c_int
    L.rt_frame_set_lane.argtypes = [vp, u32, C.POINTER(Camera), vp, u32, vp]
    L.rt_lane_count.restype = C.c_int
    L.rt_lane_count.argtypes = [vp]
    L.rt_render_region.restype = C.c_int
    L.rt_render_region.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, u32]
    L.rt_bulk_trace.restype = C.c_int
    L.rt_bulk_trace.argtypes = [vp, C.POINTER(Camera), vp, u32, vp, u32, u32, u32, u32, u32, u32, vp, u32]
    L.rt_render_region_debug.restype = C.c_int
    L.rt_render_region_debug.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, vp, vp, vp, vp, u32]
    L.rt_render_region_device.restype = C.c_int
    L.rt_render_region_device.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, vp, u32]
    L.rt_render_tiles_device.restype = C.c_int
    L.rt_render_tiles_device.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, vp, u32]
    L.rt_render_tiles_peer.restype = C.c_int
    L.rt_render_tiles_peer.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, vp, u32]
    L.rt_push_tiles_device.restype = C.c_int
    L.rt_push_tiles_device.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, vp, vp, u32]
    L.rt_render_push_tiles.restype = C.c_int
    L.rt_render_push_tiles.argtypes = [vp, C.POINTER(Camera), vp, u32, vp, u32, u32, u32, u32, u32, u32, vp, vp, vp, u32, vp, u32]
    L.rt_render_stripes_host.restype = C.c_int
    L.rt_render_stripes_host.argtypes = [vp, C.POINTER(Camera), vp, u32, vp, u32, u32, u32, u32, u32, vp, vp, vp, u32, vp, u32]
    L.rt_frame_complete_device.restype = C.c_int
    L.rt_frame_complete_device.argtypes = [vp, vp, u32, u32, u32, vp]
    L.rt_frame_release_device.restype = C.c_int
    L.rt_frame_release_device.argtypes = [vp, vp, u32, vp]
    L.rt_trace_results_max_bytes.restype = C.c_uint64
    L.rt_trace_results_max_bytes.argtypes = [C.c_uint64]
    L.rt_encode_trace_results_device.restype = C.c_int
    L.rt_encode_trace_results_device.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp, u32]
    L.rt_bulk_trace_encoded.restype = C.c_int
    L.rt_bulk_trace_encoded.argtypes = [vp, C.POINTER(Camera), vp, u32, vp, u32, u32, u32, u32, u32, u32, vp, C.c_uint64,
                                        C.POINTER(C.c_uint64), u32]
    L.rt_device_alloc.restype = C.c_int
    L.rt_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.rt_device_free.restype = C.c_int
    L.rt_device_free.argtypes = [C.c_int, vp]
    L.rt_ipc_export.restype = C.c_int
    L.rt_ipc_export.argtypes = [vp, C.c_char_p]
    L.rt_ipc_open.restype = C.c_int
    L.rt_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.rt_ipc_close.restype = C.c_int
    L.rt_ipc_close.argtypes = [C.c_int, vp]
    L.rt_untile_device.restype = C.c_int
    L.rt_untile_device.argtypes = [vp, u32, u32, u32, u32, u32, vp, vp, vp, u32]
    L.rt_measure_fma_peak.restype = C.c_int
    L.rt_measure_fma_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.rt_get_stats.restype = C.c_int
    L.rt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.rt_get_bvh_info.restype = C.c_int
    L.rt_get_bvh_info.argtypes = [vp, u32, C.POINTER(BvhInfo)]
    L.rt_synchronize.restype = C.c_int
    L.rt_synchronize.argtypes = [vp]
    L.rt_cancel.restype = C.c_int
    L.rt_cancel.argtypes = [vp]
    L.rt_host_alloc.restype = vp
    L.rt_host_alloc.argtypes = [C.c_size_t]
    L.rt_host_register.restype = C.c_int
    L.rt_host_register.argtypes = [C.c_int, vp, C.c_size_t, C.POINTER(vp)]
    L.rt_host_unregister.restype = C.c_int
    L.rt_host_unregister.argtypes = [C.c_int, vp]
    L.rt_host_free.restype = None
    L.rt_host_free.argtypes = [vp]
    L.rt_group_create.restype = C.c_int
    L.rt_group_create.argtypes = [C.POINTER(SceneDesc), vp, u32, C.POINTER(vp)]
    L.rt_group_destroy.restype = None
    L.rt_group_destroy.argtypes = [vp]
    L.rt_group_device_count.restype = C.c_int
    L.rt_group_device_count.argtypes = [vp]
    L.rt_group_scene.restype = vp
    L.rt_group_scene.argtypes = [vp, u32]
    L.rt_group_last_error.restype = C.c_char_p
    L.rt_group_last_error.argtypes = [vp]
    L.rt_group_bulk_trace.restype = C.c_int
    L.rt_group_bulk_trace.argtypes = [vp, C.POINTER(Camera), vp, u32, vp, u32, u32, u32, u32, u32, u32, vp, u32]
    _LIB = L
    return L
