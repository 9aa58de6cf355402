// rt_internal.h -- declarations shared by the translation units of librt_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_types.h"
#include "rt_bvh.cuh"

// Device-resident mesh (owned by rt_scene)
struct MeshDevice {
    RtNode   *nodes = nullptr;
    RtNode   *nodes_any = nullptr;   // any-hit hierarchy (== nodes when both leaf sizes agree)
    rt_f4    *tri32 = nullptr;
    double   *tri64 = nullptr;
    double   *fverts = nullptr;
    double   *fnormals = nullptr;
    uint32_t *fmat = nullptr;
    double   *materials = nullptr;
    uint32_t  n_faces = 0, n_nodes = 0, n_nodes_any = 0, n_materials = 0;
    int       root_ref = 0, root_ref_any = 0;
    float     lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
    RtNodeQ  *qnodes = nullptr, *qnodes_any = nullptr;      // FP32_FAST: quantised 64-byte nodes (qnodes_any == qnodes when the hierarchies agree)
    float     qorg[3] = { 0, 0, 0 }, qstep[3] = { 1, 1, 1 };
    double    vlo[3] = { 0, 0, 0 }, vhi[3] = { 0, 0, 0 };   // float64 bounds of all vertices of the mesh (rt_flatten_faces)
    rt_bvh_info info;
};

bool rt_build_bvh(MeshDevice &m, bool want32, bool want64, int max_leaf, int max_leaf_any, int sm_count, cudaStream_t st, std::string &err);
cudaError_t rt_flatten_faces(MeshDevice &m, const double *d_verts, uint32_t n_verts, const double *d_normals, uint32_t n_normals,
                             const uint32_t *d_faces, cudaStream_t st, int *bad);

// obj_ingest.cu : OBJ text (host) -> indexed mesh arrays on the device
bool rt_obj_ingest(const char *text, uint64_t n, const uint64_t *usemtl_off, const uint32_t *usemtl_mat, uint32_t n_usemtl, uint32_t default_mat,
                   cudaStream_t st, double **d_verts, uint32_t *n_verts, double **d_normals, uint32_t *n_normals, uint32_t **d_faces,
                   uint32_t *n_faces, float *ms, std::string &err);

// Per-render device scratch + launch description
struct RenderJob {
    FrameParams fp;
    const MeshView *d_meshes;
    const ObjRec   *d_objs;
    const LightRec *d_lights;
    InlineFrame fr;           // lights / object positions by value when they fit
    uint32_t precision;
    uint32_t flags;
    // scratch (device)
    HitRec   *d_hits;         // capacity = ray_count
    double   *d_hit_colour;   // 3 per hit
    uint32_t *d_hit_lit;      // per hit
    double   *d_hit_geom;     // split K4: 6 per hit (P, N) or nullptr
    bool      split_shade;
    bool      shade32;        // FP32_FAST: opt-in fp32 shading (RT_B200_SHADE32=1)
    unsigned int *d_counters; // [0]=next ray (primary), [1]=hit count, [2]=next hit (shade), [3]=next hit (write)
    Counters *d_stats;        // [0] = K2 primary, [1] = K4 shadow rays (filled with RT_FLAG_COUNT)
    // outputs (device)
    uint8_t  *d_out_rgb;
    double   *d_out_colour;   // optional dense debug outputs
    int32_t  *d_out_face;
    int32_t  *d_out_obj;
    uint32_t *d_out_lit;
    int       sm_count;
    bool      single_mesh;    // exactly one object => the persistent kernel takes its mesh by value
    MeshView  h_mesh0;        // host copy of object 0's mesh view
    double    h_obj0_pos[3];
    bool      primary_persistent;  // default primary kernel of FP32_FAST (RT_B200_PERSISTENT=1), overridable per call
    int       shade_minb;     // K4 occupancy experiment: min resident CTAs per SM (register cap)
    bool      fused;          // FP32_FAST: one fused kernel per frame instead of K2 -> K4 -> K5
    bool      sm_affine;      // ... in its persistent SM-affine form (RT_FLAG_SM_AFFINE)
    unsigned int *d_chunk_state; size_t chunk_state_words;
    unsigned int *d_tune;     // SceneView::tune (scene-level, shared by all lanes) or nullptr
    unsigned int *d_tune_host; // device address of the host-mapped copy of the bound
    bool      any_plain;      // launch the kernel variant with the plain any-hit loop (bound 1)
};

// render.cu : enqueue K2 (primary) / K4 (shadow+shade) / K5 (framebuffer) on `st`
cudaError_t rt_launch_primary(const RenderJob &job, cudaStream_t st);
cudaError_t rt_launch_shade(const RenderJob &job, cudaStream_t st);
cudaError_t rt_launch_write(const RenderJob &job, cudaStream_t st);
cudaError_t rt_launch_fused(const RenderJob &job, cudaStream_t st);
cudaError_t rt_launch_untile(const uint8_t *d_tiles, uint8_t *d_out, int W, int H, int tile_w, int tile_h, int n_ranks,
                             int tiles_per_rank_max, bool rowmajor, cudaStream_t st);

// render.cu : FMA-pipe micro-benchmark
cudaError_t rt_fma_peak(bool fp64, int sm_count, double *tflops);
cudaError_t rt_launch_clear_tiles(const FrameParams &fp, bool rowmajor, uint8_t *d_out, cudaStream_t st);
cudaError_t rt_launch_push_tiles(const FrameParams &fp, bool rowmajor, const uint8_t *d_tiles, uint8_t *d_out, const PushSync &sy,
                                 int sm_count, cudaStream_t st);
cudaError_t rt_launch_frame_complete(uint32_t *d_slot, uint32_t seq, int n_ranks, bool release, uint32_t timeout_ms, cudaStream_t st);
cudaError_t rt_launch_frame_release(uint32_t *d_slot, uint32_t seq, cudaStream_t st);
cudaError_t rt_launch_wait_consumed(uint32_t *d_slot, uint32_t seq, uint32_t rank, uint32_t timeout_ms, cudaStream_t st);
cudaError_t rt_launch_signal_arrive(uint32_t *d_slot, uint32_t seq, uint32_t rank, cudaStream_t st);

// rt_api.cu : region form of rt_render_stripes_host (used by the multi-GPU group handle, rt_group.cu)
int rt_internal_render_stripes(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                               uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t stripe, uint32_t rank, uint32_t n_ranks,
                               void *d_local_framebuffer, void *h_framebuffer, void *d_sync_slot, uint32_t seq, void *stream, uint32_t flags);
int rt_internal_fail(rt_scene *s, int code, const std::string &msg);

// wire.cu : comms.TraceResults protobuf encoder
size_t rt_wire_scratch_bytes(uint64_t n_pixels);
cudaError_t rt_launch_wire_encode(const uint8_t *d_rgb, uint64_t n_pixels, uint8_t *d_out, uint64_t capacity, uint64_t *d_len,
                                  void *d_scratch, cudaStream_t st);
