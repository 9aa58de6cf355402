/*
 * rt_b200.h -- C ABI of the B200 (sm_100a) GPU render backend.
 *
 * Drop-in boundary for the render hot path of MWindels/distributed-raytracer: the per-pixel
 * double loop in worker/distributed/main.go:67-88 (Tracer.BulkTrace) and
 * worker/sequential/main.go:21-28 (draw), i.e. every call of
 * tracer.Trace (worker/shared/tracer/tracer.go:81-91), is replaced by ONE call of
 * rt_bulk_trace / rt_render_region per screen region.  master/pool, comms.proto, registration and
 * heartbeats are untouched.  The cgo binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no torch / C++ types; every pointer argument is caller-owned and is not retained
 *    after the call returns (cgo rule: C may not keep Go pointers) -- the library copies.
 *  - every entry point returns 0 on success, a negative rt_status otherwise; no exception crosses
 *    the ABI.  rt_last_error(handle): the calling thread's failure on that handle if its last call
 *    failed (concurrent callers do not see each other's errors), else the handle's most recent
 *    message; rt_last_error(NULL): the calling thread's last call ("" after a success).
 *  - handles are internally synchronised: BulkTrace may be entered concurrently from several
 *    goroutines (final_report.pdf p.6 s2.3) and the calls then run concurrently on the GPU, one
 *    lane each (see RT_LANE below); calls may arrive on any OS thread.
 *  - there is NO CPU fallback: without a CUDA device / the sm_100a kernels every call fails with
 *    RT_ERR_CUDA.
 */
#ifndef RT_B200_H
#define RT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 2

typedef enum {
    RT_OK            = 0,
    RT_ERR_INVALID   = -1,  /* bad argument (NULL, out-of-range index, empty region ...)          */
    RT_ERR_CUDA      = -2,  /* CUDA runtime / no device / kernel image missing                    */
    RT_ERR_OOM       = -3,  /* host or device allocation failed                                   */
    RT_ERR_STATE     = -4,  /* rt_render_* before rt_frame_set, destroyed handle ...               */
    RT_ERR_CANCELLED = -5   /* rt_cancel() hit while rendering (ctx.Err(), distributed/main.go:73) */
} rt_status;

/* precision modes (BASELINE.json north_star) */
#define RT_FP64_EXACT 0u   /* mirrors the reference's float64 arithmetic order; bit-exact vs worker/sequential */
#define RT_FP32_FAST  1u   /* fp32 BVH traversal + intersection; fp64 re-intersection of the winner + fp64 shading */

/* scene-create flags */
#define RT_SCENE_DEFAULT     0u

/* render flags */
#define RT_LAYOUT_COLMAJOR   0x0u  /* out index (i*h + j)   -- BulkTrace, distributed/main.go:82            */
#define RT_LAYOUT_ROWMAJOR   0x1u  /* out index (j*w + i)   -- image order                                  */
#define RT_QUANT_RGB255      0x0u  /* uint8(255*c)          -- colour.RGB.RGB(), colour.go:59-61            */
#define RT_QUANT_RGBA16_SHR8 0x2u  /* uint32(0xFFFF*c) >> 8 -- colour.RGB.RGBA() + go-sdl2 Surface.Set      */
#define RT_FLAG_COUNT        0x4u  /* accumulate traversal counters (slower; for roofline accounting)       */
#define RT_FLAG_BRUTE_FORCE  0x8u  /* FP64_EXACT only: ignore the BVH and the object-level candidate test of tracer.go:32, test every
                                      face of every object (debug; see DESIGN.md s4 for where that differs)   */
#define RT_FLAG_SIMPLE       0x10u /* FP32_FAST only: one-thread-per-ray primary kernel (the default: fastest measured) */
#define RT_FLAG_PERSISTENT   0x20u /* FP32_FAST only: persistent warp-cooperative primary kernel (ballot/shuffle refill) */
#define RT_FLAG_FUSED        0x40u /* FP32_FAST only: ONE kernel per frame -- a warp traces its 8x4 pixel block, shades the
                                      hits (shadow rays + Phong) and writes the pixels; no hit list, no K4/K5 launches */
#define RT_FLAG_PIPELINE     0x80u /* the three-kernel pipeline K2 -> K4 -> K5 even if the scene's default is the fused kernel */
#define RT_FLAG_SM_AFFINE    0x200u /* FP32_FAST, fused frame kernel only: persistent warps, every SM works through screen-local chunks of
                                      the frame (one atomic per warp refill) so that the warps sharing an L1 walk the same part of the
                                      tree; RT_B200_SM_AFFINE=1 makes it the scene's default */
#define RT_FLAG_NO_CULL      0x100u /* trace every pixel block of the region: no projected-root-box culling (A/B, tests) */

/* Lanes.  A scene handle owns one immutable copy of the geometry (+ BVH) and several LANES -- per-call resources: stream,
 * hit-list scratch, staging buffers, frame state -- the way the reference worker serves every BulkTrace RPC on its own
 * goroutine against one shared immutable scene (worker/distributed/main.go:27-32).  Lane 0 serves the stateful two-call API
 * (rt_frame_set + rt_render_*); concurrent rt_bulk_trace calls each take a free lane from the pool (lanes >= 1) and run
 * concurrently on the GPU; the device API selects its lane with RT_LANE(k) in `flags` (default 0), at most one frame in
 * flight per lane and stream.  RT_B200_LANES (environment, default 9) sets the number of lanes. */
#define RT_LANE(k)           ((uint32_t)(k) << 24)

typedef struct rt_scene rt_scene;

/* One immutable mesh (shared/state/mesh.go:100-106, flattened).  All arrays object-space. */
typedef struct {
    const double   *verts;       /* n_verts*3                                         mesh.go:101  */
    uint32_t        n_verts;
    const double   *normals;     /* n_normals*3, already normalised (mesh.go:202); n_normals==0
                                    selects the flat-normal path (object.go:82,91-95)              */
    uint32_t        n_normals;
    const uint32_t *faces;       /* n_faces*7: v0 v1 v2 n0 n1 n2 mat, in ENUMERATION ORDER; the face
                                    index reported/tie-broken on is the position in this array
                                    (mesh.go:21-27)                           */
    uint32_t        n_faces;
    const double   *materials;   /* n_materials*10: Ka[3] Kd[3] Ks[3] Ns              mesh.go:94-97 */
    uint32_t        n_materials;
} rt_mesh_desc;

typedef struct {
    const rt_mesh_desc *meshes;
    uint32_t            n_meshes;
    const double       *obj_pos;    /* n_objects*3   state.Object.Pos, object.go:17-22               */
    const uint32_t     *obj_mesh;   /* n_objects     index into meshes                               */
    uint32_t            n_objects;
    int32_t             device;     /* CUDA ordinal                                                  */
    uint32_t            precision;  /* RT_FP64_EXACT | RT_FP32_FAST                                  */
    uint32_t            flags;      /* RT_SCENE_*                                                    */
} rt_scene_desc;

/* state.Camera (camera.go:20-24) + tan(Fov/2) computed by the host with Go's own math.Tan
 * (tracer.go:17) so the frame constant is bit-identical; pass NaN to let the library use its
 * restatement of Go's math.Tan. */
typedef struct {
    double pos[3], forward[3], left[3], up[3];
    double fov;
    double tan_half_fov;
} rt_camera;

/* state.Light (light.go:10-13); col already float64 in [0,1] (colour.NewRGB, colour.go:28-30) */
typedef struct {
    double pos[3];
    double col[3];
} rt_light;

typedef struct {
    uint64_t primary_rays;     /* = w*h of the last render                                         */
    uint64_t shadow_rays;      /* = hit pixels * lights (tracer.go:60-64)                          */
    uint64_t hit_pixels;
    uint64_t nodes_fetched;    /* BVH node records fetched, all kernels (only with RT_FLAG_COUNT)  */
    uint64_t tris_fetched;     /* triangle records fetched, all kernels (only with RT_FLAG_COUNT)  */
    uint64_t nodes_primary;    /* ... by the primary kernel (K2)                                   */
    uint64_t tris_primary;
    uint64_t nodes_shade;      /* ... by the shadow rays of the shade kernel (K4)                  */
    uint64_t tris_shade;
    uint64_t node_bytes;       /* sizeof one node record                                           */
    uint64_t tri_bytes;        /* sizeof one triangle record in the active precision               */
    float    ms_primary;       /* CUDA-event durations of the last render's kernels, on the        */
    float    ms_shade;         /* library's stream (valid after a synchronous render call)         */
    float    ms_write;
    float    ms_total;
    /* self-tuned by the kernels (FP32_FAST): the share of this handle's recent shadow rays that were OCCLUDED and the number of
     * leaves an any-hit ray currently postpones before it tests them (rt_traverse.cuh; lit scenes batch 16 or 8, 1 = the plain loop
     * for scenes whose shadow rays are mostly occluded) */
    float    occluded_fraction;
    uint32_t any_defer;
} rt_stats;

typedef struct {
    uint64_t n_faces, n_nodes, n_leaves;
    uint32_t max_depth, max_leaf;
    float    ms_build;         /* CUDA-event time of the whole on-GPU build of this mesh            */
    float    sah_cost;
    /* the any-hit (shadow-ray) hierarchy over the same triangles; equals the fields above when both leaf sizes agree */
    uint64_t n_nodes_any, n_leaves_any;
    uint32_t max_depth_any, max_leaf_any;
    /* rt_scene_create_obj only: device time of the OBJ ingest and what it found (0 otherwise) */
    float    ms_ingest;
    uint32_t n_verts_ingested, n_normals_ingested;
} rt_bvh_info;

/* One mesh as the TEXT of its Wavefront OBJ file.  The library parses it on the GPU with gwob's published
 * semantics (mesh.go:109-133,166-204): v / vn / f lines, coordinates parsed to float32 then widened (mesh.go:176), corners
 * "v", "v/t", "v//n", "v/t/n" with 1-based or negative indices, triangles and quads (quad -> (0,1,2) + (2,3,0)), faces in file
 * order; vertex normals stored normalised (mesh.go:202).  What it does NOT do is de-duplicate by value (mesh.go:146-203): that
 * changes no rendered bit.  The small MTL file stays with the host: it hands over the material table (as in rt_mesh_desc)
 * and, for every `usemtl` line of the OBJ text, its byte offset and the material index it selects. */
typedef struct {
    const char     *obj_text;         /* obj_bytes bytes, HOST memory, need not be NUL-terminated                     */
    uint64_t        obj_bytes;
    const double   *materials;        /* n_materials*10: Ka[3] Kd[3] Ks[3] Ns                       mesh.go:94-97,151-154 */
    uint32_t        n_materials;
    const uint64_t *usemtl_offset;    /* n_usemtl byte offsets of the `usemtl` lines, ascending                        */
    const uint32_t *usemtl_material;  /* n_usemtl material indices                                                    */
    uint32_t        n_usemtl;
    uint32_t        default_material; /* material of the faces above the first usemtl line          mesh.go:151       */
} rt_obj_mesh_desc;

/* ---- lifecycle (register(): distributed/main.go:101-129 => create; re-register => destroy+create) */
int  rt_scene_create(const rt_scene_desc *desc, rt_scene **out);
/* the same registration from OBJ text: parse on the GPU -> flatten -> LBVH, nothing but the file's bytes crosses PCIe */
int  rt_scene_create_obj(const rt_obj_mesh_desc *meshes, uint32_t n_meshes, const double *obj_pos, const uint32_t *obj_mesh,
                         uint32_t n_objects, int32_t device, uint32_t precision, rt_scene **out);
void rt_scene_destroy(rt_scene *scene);
const char *rt_last_error(const rt_scene *scene);
int  rt_abi_version(void);
int  rt_device_count(void);

/* ---- per-frame mutable state (the gob `diff`: distributed/main.go:56-64, environment.go:65-69) --
 * obj_pos, wherever it appears below, must hold exactly 3 * n_objects doubles (n_objects of rt_scene_create): like a Go
 * slice header passed through cgo as a bare pointer, the ABI carries no length for it.  An object whose position is NaN
 * (its x component) is ABSENT from the order and can never be hit: the reference rebuilds the object R-tree from the diff
 * alone (environment.go:73-98), so a registered object that an order does not list is simply not rendered.  n_lights is
 * unbounded; the lit mask of the debug outputs has one bit per light modulo 32. */
int rt_frame_set(rt_scene *scene, const rt_camera *cam, const rt_light *lights, uint32_t n_lights,
                 const double *obj_pos /* n_objects*3 or NULL = unchanged */);
/* the same for an explicit lane of the device API */
int rt_frame_set_lane(rt_scene *scene, uint32_t lane, const rt_camera *cam, const rt_light *lights, uint32_t n_lights,
                      const double *obj_pos);
int rt_lane_count(rt_scene *scene);

/* ---- render one screen region (distributed/main.go:67-88; sequential/main.go:21-28) ------------
 * Pixel (x+i, y+j), i<w, j<h, of a screen_w x screen_h screen (tracer.go:81 takes the FULL screen
 * size).  out_rgb: w*h*3 bytes, HOST memory (pageable or pinned), misses are (0,0,0).  Blocks until
 * the bytes are in out_rgb. */
int rt_render_region(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y,
                     uint32_t w, uint32_t h, uint8_t *out_rgb, uint32_t flags);

/* frame state + render of one WorkOrder = what one BulkTrace RPC does.  Thread-safe and CONCURRENT: every call takes its own
 * lane, so several goroutines can have orders in flight on one registered scene (obj_pos NULL = the registered positions). */
int rt_bulk_trace(rt_scene *scene, const rt_camera *cam, const rt_light *lights, uint32_t n_lights,
                  const double *obj_pos, uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y,
                  uint32_t w, uint32_t h, uint8_t *out_rgb, uint32_t flags);

/* Parity/debug variant: any of the outputs may be NULL.  out_colour: w*h*3 float64 (the value
 * tracer.Trace returns), out_face: w*h int32 face index of the primary hit (-1 = miss), out_obj:
 * w*h int32 object index, out_lit: w*h uint32 bit k = light k contributed (tracer.go:64). */
int rt_render_region_debug(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y,
                           uint32_t w, uint32_t h, uint8_t *out_rgb, double *out_colour, int32_t *out_face,
                           int32_t *out_obj, uint32_t *out_lit, uint32_t flags);

/* Device-resident variant: d_out_rgb is DEVICE memory on the scene's device (w*h*3 bytes); the work
 * is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the library's own stream) and the
 * call returns without synchronising when stream != NULL. */
int rt_render_region_device(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y,
                            uint32_t w, uint32_t h, void *d_out_rgb, void *stream, uint32_t flags);

/* Multi-GPU screen-tile sharding (master/main.go:54-91 partition(), re-designed as interleaved
 * tiles): the full screen is cut into tile_w x tile_h tiles; the tile at column X, row Y has id
 * t = Y*tiles_x + (X - 3*Y) mod tiles_x (each tile row rotated by 3 more positions, so that the
 * owners form diagonals, not column stripes, when n_ranks divides tiles_x); this call renders the
 * tiles t with t % n_ranks == rank and packs them, in increasing t, into d_out_tiles (DEVICE,
 * n_local_tiles * tile_w*tile_h*3 bytes, each tile row-major, edge tiles padded with zeros). */
int rt_render_tiles_device(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t tile_w,
                           uint32_t tile_h, uint32_t rank, uint32_t n_ranks, void *d_out_tiles,
                           void *stream, uint32_t flags);
/* Peer-memory variant (the multi-GPU path of choice): renders the same tiles, but the framebuffer writer stores each hit
 * pixel -- and the clear of this rank's tiles -- straight into the FULL screen_w*screen_h*3 framebuffer d_framebuffer in
 * `flags` layout.  d_framebuffer may live on another GPU of the box (rt_ipc_open of rank 0's buffer): the stores travel
 * over NVLink as the kernel runs, so there is no gather and no de-interleave; the ranks only need a barrier afterwards.
 * Pixels of other ranks' tiles are never touched. */
int rt_render_tiles_peer(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t tile_w, uint32_t tile_h,
                         uint32_t rank, uint32_t n_ranks, void *d_framebuffer, void *stream, uint32_t flags);
/* Peer-memory push (preferred across GPUs): copies this rank's packed tiles (output of rt_render_tiles_device, local) into
 * the FULL framebuffer d_framebuffer in `flags` layout with 16-byte coalesced stores.  d_framebuffer normally is rank 0's
 * buffer opened with rt_ipc_open, so the stores travel over NVLink; no gather, no de-interleave on rank 0. */
int rt_push_tiles_device(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t tile_w, uint32_t tile_h,
                         uint32_t rank, uint32_t n_ranks, const void *d_tiles, void *d_framebuffer, void *stream,
                         uint32_t flags);
/* rt_frame_set + rt_render_tiles_device + rt_push_tiles_device in ONE call (the per-frame call of a rank in the multi-GPU
 * pipeline: at 8 GPUs a frame is ~0.12 ms, so host crossings matter).  Asynchronous: `stream` must not be NULL.
 *
 * Frame hand-off without a collective: d_sync_slot (may be NULL = no hand-off) points at the RT_SYNC_SLOT_BYTES record of
 * the framebuffer slot, in the gathering rank's memory like the framebuffer itself (zeroed once; opened with rt_ipc_open
 * by the others).  Frames of a slot are numbered seq = 1, 2, ...  The push kernel first waits -- on the device -- until the
 * gathering rank has released frame seq-1 of the slot, stores its tiles over NVLink, and its last CTA then publishes
 * arrive[rank] = seq with a system-scope release.  A wait gives up after RT_B200_SYNC_TIMEOUT_MS (2000) and records
 * 0x200+rank (push) or 0x100+rank (complete) in the record's error word instead of hanging the GPU. */
#define RT_SYNC_SLOT_BYTES 256u   /* uint32 arrive[32]; uint32 consumed; uint32 error; padding */
int rt_render_push_tiles(rt_scene *scene, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos,
                         uint32_t screen_w, uint32_t screen_h, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                         void *d_tiles, void *d_framebuffer, void *d_sync_slot, uint32_t seq, void *stream, uint32_t flags);
/* The multi-GPU END-TO-END path: the frame lands in HOST memory over every GPU's own PCIe link.  The screen is cut into
 * stripes that are contiguous in the output layout (column-major BulkTrace layout: `stripe` pixel columns x full height,
 * `stripe` a multiple of 8; RT_LAYOUT_ROWMAJOR: `stripe` pixel rows x full width, a multiple of 4), stripe t owned by rank
 * t % n_ranks.  One call = rt_frame_set + render of the rank's stripes in place into d_local_framebuffer (DEVICE,
 * screen_w*screen_h*3, final layout) + ONE strided copy-engine transfer of those stripes into h_framebuffer (the address
 * rt_host_register returned for the shared host framebuffer).  d_sync_slot / seq as in rt_render_push_tiles (the record
 * may live in registered host memory too): waits for the slot's release before the transfer, publishes arrive[rank]
 * after it.  Asynchronous on `stream`. */
int rt_render_stripes_host(rt_scene *scene, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos,
                           uint32_t screen_w, uint32_t screen_h, uint32_t stripe, uint32_t rank, uint32_t n_ranks,
                           void *d_local_framebuffer, void *h_framebuffer, void *d_sync_slot, uint32_t seq, void *stream,
                           uint32_t flags);
/* Gathering rank, in stream order on `stream`: waits (on the device) until all n_ranks ranks have pushed frame `seq` of the
 * slot; everything enqueued on `stream` afterwards sees the complete frame.  release != 0 hands the slot back at once;
 * otherwise call rt_frame_release_device after the consumer (e.g. the device->host copy) on the same stream. */
int rt_frame_complete_device(rt_scene *scene, void *d_sync_slot, uint32_t seq, uint32_t n_ranks, uint32_t release, void *stream);
int rt_frame_release_device(rt_scene *scene, void *d_sync_slot, uint32_t seq, void *stream);
/* Device buffers shareable between the one-process-per-GPU ranks of a box (CUDA IPC; 64-byte opaque handles) */
int rt_device_alloc(int device, size_t bytes, void **out);
int rt_device_free(int device, void *p);
int rt_ipc_export(void *d_ptr, unsigned char handle[64]);
int rt_ipc_open(int device, const unsigned char handle[64], void **out);
int rt_ipc_close(int device, void *p);
/* Inverse of the packing above, run on the gathering rank: d_all_tiles holds the n_ranks packed
 * buffers back to back (each padded to tiles_per_rank_max tiles); writes the screen_w*screen_h*3
 * framebuffer d_out_rgb in `flags` layout. */
int rt_untile_device(rt_scene *scene, uint32_t screen_w, uint32_t screen_h, uint32_t tile_w, uint32_t tile_h,
                     uint32_t n_ranks, const void *d_all_tiles, void *d_out_rgb, void *stream, uint32_t flags);

/* ---- one worker process, several GPUs (master/pool dispatches ONE BulkTrace per worker: pool.go:148-197 ->
 * worker/distributed/main.go:46-91) -----------------------------------------------------------------------------------------
 * rt_group_create registers the scene on every listed device (devices == NULL or n_devices == 0: every GPU of the box); it is
 * register() (distributed/main.go:101-129) for a worker that owns the whole box.  rt_group_bulk_trace answers ONE order with all
 * of them: the region is cut into stripes that are contiguous in the reply's layout (8 pixel columns x region height in
 * BulkTrace's layout, 4 pixel rows x region width with RT_LAYOUT_ROWMAJOR), stripe t is rendered by device t mod N, and every
 * device moves its stripes with one strided copy over its own PCIe link into out_rgb (directly if out_rgb is pinned memory from
 * rt_host_alloc, else through a pinned staging frame).  No collective, no torch.distributed, no second process: the calling
 * thread enqueues N command sequences and waits for N streams.  Same arguments, result bytes and error behaviour as
 * rt_bulk_trace on one device (the frame is byte-identical); thread-safe and concurrent (RT_B200_GROUP_LANES calls in flight,
 * default 8).  A device may be listed more than once (it then holds that many copies of the scene and renders that many
 * shares of the stripes).  The per-device scene handles stay owned by the group (rt_group_scene: introspection only). */
typedef struct rt_group rt_group;
int  rt_group_create(const rt_scene_desc *desc /* .device ignored */, const int32_t *devices, uint32_t n_devices, rt_group **out);
void rt_group_destroy(rt_group *group);
int  rt_group_device_count(const rt_group *group);
rt_scene *rt_group_scene(rt_group *group, uint32_t k);
const char *rt_group_last_error(const rt_group *group);
int  rt_group_bulk_trace(rt_group *group, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos,
                         uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                         uint32_t flags);

/* ---- result wire format --------------------------------------------------------
 * The reference worker returns a region as the protobuf message comms.TraceResults
 * (shared/comms/comms.proto:34-41: repeated Colour{uint32 r=1, g=2, b=3}), one heap-allocated Colour
 * per pixel (worker/distributed/main.go:52-54,79-86), marshalled by grpc-go.  These calls produce
 * exactly the bytes proto.Marshal emits for that message -- per pixel 0x0A len [0x08 r] [0x10 g]
 * [0x18 b], zero-valued fields omitted (proto3), values >= 128 as 2-byte varints -- on the GPU, so a
 * worker hands grpc a ready payload and an unmodified master decodes it. */
uint64_t rt_trace_results_max_bytes(uint64_t n_pixels);            /* 11 * n_pixels */
/* d_rgb: n_pixels*3 bytes in BulkTrace order (DEVICE, 4-byte aligned); d_out: `capacity` bytes
 * (DEVICE); *d_len (DEVICE, 8 bytes) receives the message length -- if it exceeds `capacity` the
 * output is incomplete.  Asynchronous on `stream` (NULL: library stream, synchronous).  The lane
 * (RT_LANE(k) in flags) owns the scan scratch: one encode at a time per lane. */
int rt_encode_trace_results_device(rt_scene *scene, const void *d_rgb, uint64_t n_pixels, void *d_out,
                                   uint64_t capacity, uint64_t *d_len, void *stream, uint32_t flags);
/* rt_bulk_trace down to the wire bytes: frame state + render + encode on the device, one copy of
 * exactly *out_len bytes into `out` (HOST, `capacity` bytes; rt_host_alloc memory avoids a staging
 * copy).  RT_ERR_INVALID with *out_len set if `capacity` is too small.  Concurrent calls on one
 * handle run concurrently (one lane each), like rt_bulk_trace. */
int rt_bulk_trace_encoded(rt_scene *scene, const rt_camera *cam, const rt_light *lights, uint32_t n_lights,
                          const double *obj_pos, uint32_t screen_w, uint32_t screen_h, uint32_t x, uint32_t y,
                          uint32_t w, uint32_t h, uint8_t *out, uint64_t capacity, uint64_t *out_len,
                          uint32_t flags);

/* ---- introspection -------------------------------------------------------------------------- */
int rt_get_stats(rt_scene *scene, rt_stats *out);
int rt_get_bvh_info(rt_scene *scene, uint32_t mesh, rt_bvh_info *out);
int rt_synchronize(rt_scene *scene);
/* ctx.Err() of the reference loop (distributed/main.go:73): stops the renders that are IN FLIGHT on this handle when it is
 * called -- their kernels poll a cancel generation and give up, the blocking calls return RT_ERR_CANCELLED (their output
 * buffers then hold no frame).  Calls that start afterwards are not affected, and with nothing in flight it is a no-op:
 * nothing needs to be reset. */
int rt_cancel(rt_scene *scene);

/* Register-resident FMA micro-benchmark (roofline denominator of the flop side): runs `iters`
 * dependent-chain FMAs per thread on a full grid and returns the measured TFLOP/s (2 flops per FMA).
 * fp64 != 0 selects double precision. */
int rt_measure_fma_peak(int device, int fp64, double *out_tflops);

/* pinned host buffers so out_rgb can be DMA'd without a staging copy (optional) */
void *rt_host_alloc(size_t bytes);
void  rt_host_free(void *p);
/* Pins caller-owned HOST memory -- typically a POSIX shared-memory mapping that every per-GPU process of
 * the box has mapped -- for `device` and returns in *d_ptr the address its kernels use for it.  A framebuffer
 * and its hand-off record registered this way can be the d_framebuffer / d_sync_slot of
 * rt_render_push_tiles: every GPU then stores its tiles straight into the host framebuffer over ITS OWN
 * PCIe link (the device->host copy of a 4K frame through one link is what bounds e2e at N > 1). */
int rt_host_register(int device, void *ptr, size_t bytes, void **d_ptr);
int rt_host_unregister(int device, void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
