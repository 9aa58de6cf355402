// rt_api.cu -- implementation of the C ABI declared in include/rt_b200.h.
//
// Lifecycle mirrors the reference worker (paths under /root/reference/):
//   rt_scene_create   <- register():        worker/distributed/main.go:101-129 (immutable meshes arrive once)
//   rt_frame_set      <- BulkTrace diff:    worker/distributed/main.go:56-64  (camera, lights, object positions)
//   rt_render_region  <- BulkTrace loop:    worker/distributed/main.go:67-88  /  draw(): worker/sequential/main.go:21-28
//   rt_bulk_trace     <- one BulkTrace RPC: both of the above for one WorkOrder, concurrently callable
// No CPU fallback exists: every path below needs a CUDA device and the sm_100a kernels.
#include "rt_internal.h"
#include "go_math.h"
#include "rt_cull.h"
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <new>

// Per-call mutable resources.  The reference worker serves every BulkTrace RPC on its own goroutine against ONE immutable
// scene (worker/distributed/main.go:27-32: "No lock here because we never mutate this data"); a Lane is what one such call
// needs on the GPU: a stream, timing events, the hit-list scratch, staging buffers and its own copy of the frame state.
// Lane 0 belongs to the stateful two-call API (rt_frame_set + rt_render_*) and is the default of the device API; lanes >= 1
// form the pool that concurrent rt_bulk_trace calls draw from.  Scratch is allocated on a lane's first use.
struct Lane {
    std::mutex mu;
    bool busy = false;                 // pool bookkeeping (guarded by rt_scene::mu)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    // frame state of the call in flight (environment.go:65-69)
    FrameParams frame;
    InlineFrame inline_frame;
    bool have_frame = false;
    std::vector<ObjRec> h_objs;
    ObjRec *d_objs = nullptr;
    LightRec *d_lights = nullptr;
    uint32_t cap_lights = 0;
    // scratch
    HitRec *d_hits = nullptr; double *d_hit_colour = nullptr; uint32_t *d_hit_lit = nullptr; double *d_hit_geom = nullptr;
    size_t cap_rays = 0;
    unsigned int *d_counters = nullptr; Counters *d_stats = nullptr;
    unsigned int *d_chunk_state = nullptr; size_t cap_chunk_state = 0;   // k_frame_fused_sm
    uint8_t *d_rgb = nullptr; size_t cap_rgb = 0;
    double *d_dbg_colour = nullptr; int32_t *d_dbg_face = nullptr, *d_dbg_obj = nullptr; uint32_t *d_dbg_lit = nullptr;
    size_t cap_dbg = 0;
    uint8_t *h_pinned = nullptr; size_t cap_pinned = 0;
    // rt_bulk_trace in strips: the device->host copy of strip k runs on copy_stream while strip k+1 renders on `stream`
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaStream_t> strip_stream;          // one per strip, descending priority: strip 0's blocks are scheduled first
    std::vector<cudaEvent_t> strip_ev, copy_ev;
    cudaEvent_t start_ev = nullptr;
    // result wire format (wire.cu): encoded message, block-scan scratch, message length (device + pinned host)
    uint8_t *d_wire = nullptr; size_t cap_wire = 0;
    void *d_wire_scratch = nullptr; size_t cap_wire_scratch = 0;
    uint64_t *d_wire_len = nullptr, *h_wire_len = nullptr;
    cudaStream_t last_stream = nullptr;
    bool stats_pending = false;
    bool stats_counted = false;        // the frame whose stats are pending ran with RT_FLAG_COUNT
    rt_stats stats;
};

struct rt_scene {
    std::mutex mu;                     // lane pool + error string
    std::condition_variable cv;
    std::string err;
    int device = 0;
    uint32_t precision = RT_FP32_FAST;
    int sm_count = 148;
    int max_leaf = 4;
    int max_leaf_any = 4;              // leaf size of the any-hit (shadow-ray) hierarchy, RT_B200_MAX_LEAF_SHADE
    bool fused = false;                // FP32_FAST default: one fused kernel per frame (RT_B200_FUSED)
    bool cull = true;                  // projected-root-box culling of primary rays (RT_B200_CULL)
    int shade_minb = 8;
    bool primary_persistent = false;
    bool sm_affine = false;
    int bulk_strips = 4;
    unsigned int *h_tune = nullptr, *d_tune_host = nullptr;   // host-mapped copy of the bound (pinned; the kernels write it, the launcher reads it)
    int any_defer_fixed = 0;           // RT_B200_ANY_DEFER (0 = self-tuned)
    unsigned int *d_tune = nullptr;    // SceneView::tune: {shadow rays, occluded, postponement bound, spare}; nullptr = fixed bound (RT_B200_ANY_DEFER)
    bool split_shade = false;
    bool shade32 = false;              // RT_B200_SHADE32=1: FP32_FAST shades in fp32 too (opt-in)
    uint32_t sync_timeout_ms = 2000;   // device-side wait on a peer's hand-off flag gives up after this (master/main.go:35: 2 s per order)

    // immutable after rt_scene_create
    std::vector<MeshDevice> meshes;
    std::vector<MeshView> h_views;
    MeshView *d_meshes = nullptr;
    std::vector<ObjRec> h_objs;        // initial object positions (the lanes keep per-frame copies)

    std::vector<std::unique_ptr<Lane>> lanes;
    std::atomic<int> last_lane{ 0 };   // lane of the most recent call (rt_get_stats)
    // Cancellation (ctx.Err(), distributed/main.go:73): a generation counter.  Every render captures the generation when it
    // starts; rt_cancel bumps it and writes the new value into word RT_CNT_CANCEL of every lane's counter block, which the
    // kernels poll.  A call is cancelled iff the generation moved while it was in flight: calls that start later are unaffected,
    // nothing is ever reset.
    std::atomic<unsigned int> cancel_gen{ 0 };
    cudaStream_t cancel_stream = nullptr;
    unsigned int *h_cancel_word = nullptr;   // pinned staging word of rt_cancel
};

namespace {

// Error text.  Every entry point clears the calling thread's message on entry, so rt_last_error() describes the thread's LAST
// call: with a handle it returns this thread's failure on that handle if there is one (concurrent callers do not see each
// other's errors), else the handle's most recent message; with NULL the thread's message (handle-less failures).
thread_local std::string g_err;
thread_local const rt_scene *g_err_scene = nullptr;
thread_local std::string g_err_ret;

int fail(rt_scene *s, int code, const std::string &msg) {
    if (s) { std::lock_guard<std::mutex> lk(s->mu); s->err = msg; }
    g_err = msg;
    g_err_scene = s;
    return code;
}
struct ApiEnter { ApiEnter() { g_err.clear(); g_err_scene = nullptr; } };
#define API_ENTER() ApiEnter api_enter_guard_
#define CU(s, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail((s), e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T> cudaError_t upload(T **d, const T *h, size_t n, cudaStream_t st) {
    cudaError_t e = cudaMalloc((void **)d, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(*d, h, sizeof(T) * n, cudaMemcpyHostToDevice, st);
    return e;
}

void free_mesh(MeshDevice &m) {
    if (m.nodes_any != m.nodes) cudaFree(m.nodes_any);
    if (m.qnodes_any != m.qnodes) cudaFree(m.qnodes_any);
    cudaFree(m.qnodes);
    cudaFree(m.nodes); cudaFree(m.tri32); cudaFree(m.tri64); cudaFree(m.fverts); cudaFree(m.fnormals); cudaFree(m.fmat);
    cudaFree(m.materials);
    m = MeshDevice();
}

void destroy_lane(Lane &L) {
    if (L.stream) cudaStreamSynchronize(L.stream);
    cudaFree(L.d_objs); cudaFree(L.d_lights);
    cudaFree(L.d_hits); cudaFree(L.d_hit_colour); cudaFree(L.d_hit_lit); cudaFree(L.d_hit_geom); cudaFree(L.d_counters); cudaFree(L.d_stats);
    cudaFree(L.d_chunk_state);
    cudaFree(L.d_rgb); cudaFree(L.d_dbg_colour); cudaFree(L.d_dbg_face); cudaFree(L.d_dbg_obj); cudaFree(L.d_dbg_lit);
    if (L.h_pinned) cudaFreeHost(L.h_pinned);
    for (cudaEvent_t e : L.strip_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : L.copy_ev) cudaEventDestroy(e);
    for (cudaStream_t t : L.strip_stream) cudaStreamDestroy(t);
    if (L.start_ev) cudaEventDestroy(L.start_ev);
    if (L.copy_stream) cudaStreamDestroy(L.copy_stream);
    cudaFree(L.d_wire); cudaFree(L.d_wire_scratch); cudaFree(L.d_wire_len);
    if (L.h_wire_len) cudaFreeHost(L.h_wire_len);
    for (auto &e : L.ev) if (e) cudaEventDestroy(e);
    if (L.stream) cudaStreamDestroy(L.stream);
}

void destroy_all(rt_scene *s) {
    cudaSetDevice(s->device);
    for (auto &lp : s->lanes) destroy_lane(*lp);
    s->lanes.clear();
    for (auto &m : s->meshes) free_mesh(m);
    cudaFree(s->d_meshes);
    cudaFree(s->d_tune); s->d_tune = nullptr;
    if (s->h_tune) { cudaFreeHost(s->h_tune); s->h_tune = nullptr; s->d_tune_host = nullptr; }
    s->d_meshes = nullptr;
    if (s->cancel_stream) { cudaStreamDestroy(s->cancel_stream); s->cancel_stream = nullptr; }
    if (s->h_cancel_word) { cudaFreeHost(s->h_cancel_word); s->h_cancel_word = nullptr; }
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// pointer / count sanity of one mesh description (the indices themselves are validated on the device, k_flatten_faces)
int check_mesh_desc(const rt_mesh_desc &d, std::string &err) {
    if (d.n_faces && (!d.faces || !d.verts)) { err = "mesh: faces/verts pointer is NULL"; return RT_ERR_INVALID; }
    if (d.n_materials == 0 && d.n_faces) { err = "mesh: no materials"; return RT_ERR_INVALID; }
    if (d.n_materials && !d.materials) { err = "mesh: materials pointer is NULL"; return RT_ERR_INVALID; }
    if (d.n_normals && !d.normals) { err = "mesh: normals pointer is NULL"; return RT_ERR_INVALID; }
    return RT_OK;
}

// ---- lane pool ---------------------------------------------------------------------------------
// a free pool lane for a concurrent rt_bulk_trace (lanes >= 1; lane 0 if the scene has a single lane)
Lane *acquire_lane(rt_scene *s, int *index) {
    std::unique_lock<std::mutex> lk(s->mu);
    const int n = (int)s->lanes.size();
    for (;;) {
        for (int k = (n > 1 ? 1 : 0); k < n; k++)
            if (!s->lanes[k]->busy) { s->lanes[k]->busy = true; *index = k; return s->lanes[k].get(); }
        s->cv.wait(lk);
    }
}
void release_lane(rt_scene *s, Lane *L) {
    { std::lock_guard<std::mutex> lk(s->mu); L->busy = false; }
    s->cv.notify_one();
}
// lane of a device-API call: bits 24..31 of `flags` (RT_LANE(k)); nullptr + error if out of range
Lane *device_lane(rt_scene *s, uint32_t flags, int *index) {
    int k = (int)(flags >> 24);
    if (k >= (int)s->lanes.size()) { fail(s, RT_ERR_INVALID, "lane index (flags bits 24..31) out of range"); return nullptr; }
    *index = k;
    return s->lanes[k].get();
}

int ensure_scratch(rt_scene *s, Lane &L, size_t nrays, size_t out_bytes, bool need_rgb, bool dbg, size_t npix_out) {
    if (!L.d_counters) {
        // [0..3] per-frame counters, [RT_CNT_CANCEL] cancel generation (rt_cancel), [RT_CNT_PUSH_DONE] CTAs-done counter of k_push_tiles
        unsigned int *cnt = nullptr;
        CU(s, cudaMalloc((void **)&cnt, 8 * sizeof(unsigned int)));
        CU(s, cudaMemset(cnt, 0, 8 * sizeof(unsigned int)));
        { std::lock_guard<std::mutex> lk(s->mu); L.d_counters = cnt; }           // rt_cancel walks the lanes under s->mu
        CU(s, cudaMalloc((void **)&L.d_stats, 2 * sizeof(Counters)));
    }
    if (nrays > L.cap_rays) {
        cudaFree(L.d_hits); cudaFree(L.d_hit_colour); cudaFree(L.d_hit_lit); cudaFree(L.d_hit_geom);
        L.d_hits = nullptr; L.d_hit_colour = nullptr; L.d_hit_lit = nullptr; L.d_hit_geom = nullptr; L.cap_rays = 0;
        CU(s, cudaMalloc((void **)&L.d_hits, sizeof(HitRec) * nrays));
        CU(s, cudaMalloc((void **)&L.d_hit_colour, sizeof(double) * 3 * nrays));
        CU(s, cudaMalloc((void **)&L.d_hit_lit, sizeof(uint32_t) * nrays));
        if (s->split_shade) CU(s, cudaMalloc((void **)&L.d_hit_geom, sizeof(double) * 6 * nrays));
        L.cap_rays = nrays;
    }
    if (need_rgb && out_bytes > L.cap_rgb) {
        cudaFree(L.d_rgb); L.d_rgb = nullptr; L.cap_rgb = 0;
        CU(s, cudaMalloc((void **)&L.d_rgb, out_bytes));
        L.cap_rgb = out_bytes;
    }
    if (dbg && npix_out > L.cap_dbg) {
        cudaFree(L.d_dbg_colour); cudaFree(L.d_dbg_face); cudaFree(L.d_dbg_obj); cudaFree(L.d_dbg_lit);
        L.d_dbg_colour = nullptr; L.d_dbg_face = nullptr; L.d_dbg_obj = nullptr; L.d_dbg_lit = nullptr; L.cap_dbg = 0;
        CU(s, cudaMalloc((void **)&L.d_dbg_colour, sizeof(double) * 3 * npix_out));
        CU(s, cudaMalloc((void **)&L.d_dbg_face, sizeof(int32_t) * npix_out));
        CU(s, cudaMalloc((void **)&L.d_dbg_obj, sizeof(int32_t) * npix_out));
        CU(s, cudaMalloc((void **)&L.d_dbg_lit, sizeof(uint32_t) * npix_out));
        L.cap_dbg = npix_out;
    }
    return RT_OK;
}

// frame state of one call into lane L (caller holds L.mu)
int frame_set_lane(rt_scene *s, Lane &L, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos) {
    if (!cam) return fail(s, RT_ERR_INVALID, "rt_frame_set: camera is NULL");
    if (n_lights && !lights) return fail(s, RT_ERR_INVALID, "rt_frame_set: lights is NULL");
    CU(s, cudaSetDevice(s->device));
    FrameParams &fp = L.frame;
    for (int k = 0; k < 3; k++) { fp.cam_pos[k] = cam->pos[k]; fp.fwd[k] = cam->forward[k]; fp.left[k] = cam->left[k]; fp.up[k] = cam->up[k]; }
    fp.tan_half = std::isnan(cam->tan_half_fov) ? rt_go_tan(cam->fov / 2.0) : cam->tan_half_fov;   // tracer.go:17
    fp.n_lights = (int)n_lights;
    fp.n_obj = (int)L.h_objs.size();
    if (obj_pos)
        for (size_t k = 0; k < L.h_objs.size(); k++)
            for (int c = 0; c < 3; c++) L.h_objs[k].pos[c] = obj_pos[3 * k + c];
    InlineFrame &fr = L.inline_frame;
    fr.use_lights = n_lights <= RT_INLINE_LIGHTS;
    fr.use_objs = L.h_objs.size() <= RT_INLINE_OBJS;
    static_assert(sizeof(LightRec) == sizeof(rt_light), "light layout");
    if (fr.use_lights && n_lights) memcpy(fr.lights, lights, sizeof(LightRec) * n_lights);
    if (fr.use_objs && !L.h_objs.empty()) memcpy(fr.objs, L.h_objs.data(), sizeof(ObjRec) * L.h_objs.size());
    if (!fr.use_lights || !fr.use_objs) {
        // large light / object sets live in device memory: the copies must not overtake a render still in flight
        CU(s, cudaStreamSynchronize(L.stream));
        if (L.last_stream) CU(s, cudaStreamSynchronize(L.last_stream));
        if (!fr.use_lights) {
            if (n_lights > L.cap_lights) {
                cudaFree(L.d_lights); L.d_lights = nullptr; L.cap_lights = 0;
                CU(s, cudaMalloc((void **)&L.d_lights, sizeof(LightRec) * n_lights));
                L.cap_lights = n_lights;
            }
            CU(s, cudaMemcpyAsync(L.d_lights, lights, sizeof(LightRec) * n_lights, cudaMemcpyHostToDevice, L.stream));
        }
        if (!fr.use_objs) {
            if (!L.d_objs) CU(s, cudaMalloc((void **)&L.d_objs, sizeof(ObjRec) * L.h_objs.size()));
            CU(s, cudaMemcpyAsync(L.d_objs, L.h_objs.data(), sizeof(ObjRec) * L.h_objs.size(), cudaMemcpyHostToDevice, L.stream));
        }
        CU(s, cudaStreamSynchronize(L.stream));   // `lights` is caller memory: do not retain it
    }
    L.have_frame = true;
    return RT_OK;
}

struct RenderOut {
    uint8_t *h_rgb = nullptr; double *h_colour = nullptr; int32_t *h_face = nullptr, *h_obj = nullptr; uint32_t *h_lit = nullptr;
    uint8_t *d_rgb = nullptr;      // external device output (device variants)
    cudaStream_t ext_stream = nullptr; bool async = false;
    // a frame rendered as several launches (strips of one order): only the first zeroes the lane's counters and opens the frame's
    // timing events, only the last closes them
    bool first = true, last = true;
};

void collect_stats(rt_scene *s, Lane &L) {
    if (!L.stats_pending) return;
    unsigned int counters[4] = { 0, 0, 0, 0 };
    cudaMemcpy(counters, L.d_counters, sizeof counters, cudaMemcpyDeviceToHost);
    Counters c[2];
    memset(c, 0, sizeof c);
    if (L.stats_counted) cudaMemcpy(c, L.d_stats, sizeof c, cudaMemcpyDeviceToHost);   // traversal counters exist only for RT_FLAG_COUNT frames
    L.stats.hit_pixels = counters[1];
    L.stats.shadow_rays = (uint64_t)counters[1] * (uint64_t)L.frame.n_lights;
    L.stats.nodes_primary = c[0].nodes; L.stats.tris_primary = c[0].tris;
    L.stats.nodes_shade = c[1].nodes; L.stats.tris_shade = c[1].tris;
    L.stats.nodes_fetched = c[0].nodes + c[1].nodes;
    L.stats.tris_fetched = c[0].tris + c[1].tris;
    L.stats.node_bytes = RT_NODE_BYTES;
    L.stats.tri_bytes = s->precision == RT_FP64_EXACT ? RT_TRI64_BYTES : RT_TRI32_BYTES;
    cudaEventElapsedTime(&L.stats.ms_primary, L.ev[0], L.ev[1]);
    cudaEventElapsedTime(&L.stats.ms_shade, L.ev[1], L.ev[2]);
    cudaEventElapsedTime(&L.stats.ms_write, L.ev[2], L.ev[3]);
    cudaEventElapsedTime(&L.stats.ms_total, L.ev[0], L.ev[3]);
    L.stats.occluded_fraction = 0.f; L.stats.any_defer = s->h_views.empty() ? 0u : s->h_views[0].any_defer;
    if (s->d_tune) {
        unsigned int t[4] = { 0, 0, 0, 0 };
        cudaMemcpy(t, s->d_tune, sizeof t, cudaMemcpyDeviceToHost);
        L.stats.occluded_fraction = t[0] ? (float)t[1] / (float)t[0] : 0.f;
        L.stats.any_defer = t[2];
    }
    L.stats_pending = false;
}

// one fused kernel per frame (FP32_FAST): the scene's default, RT_FLAG_FUSED / RT_FLAG_PIPELINE per call
static bool will_fuse(const rt_scene *s, uint32_t flags) {
    return s->precision == RT_FP32_FAST && !s->shade32 && !(flags & (RT_FLAG_PIPELINE | RT_FLAG_PERSISTENT)) && (s->fused || (flags & RT_FLAG_FUSED));
}

// the whole render path on one lane: clear -> K2 -> K4 -> K5 -> (copy out).  Caller holds L.mu.
int render_lane(rt_scene *s, Lane &L, const FrameParams &fp_in, uint32_t flags, size_t npix_out, const RenderOut &out) {
    if (!L.have_frame) return fail(s, RT_ERR_STATE, "render before rt_frame_set");
    CU(s, cudaSetDevice(s->device));
    const unsigned int my_gen = s->cancel_gen.load();     // rt_cancel during this call bumps it (stats of earlier frames are collected lazily)
    FrameParams fp = fp_in;
    fp.cancel_gen = my_gen;
    const size_t nrays = ray_count(fp);
    const size_t out_bytes = npix_out * 3;
    const bool dbg = out.h_colour || out.h_face || out.h_obj || out.h_lit;
    const bool internal_rgb = out.d_rgb == nullptr;
    const bool fused = will_fuse(s, flags);
    // The hit-list scratch is sized for the UN-culled ray count of the region / tile set, so that a camera move (another cull
    // rectangle) never re-allocates: cudaMalloc / cudaFree wait for the device, and in the multi-GPU pipeline a peer's hand-off
    // kernel may be spinning on this very rank's progress.  The fused kernel needs no hit list at all.
    size_t full_rays = fp.tile_w > 0 ? (size_t)fp.n_local_tiles * (size_t)(fp.tile_w * fp.tile_h)
                                     : (size_t)((fp.w + 7) / 8) * (size_t)((fp.h + 3) / 4) * 32u;
    int rc = ensure_scratch(s, L, fused ? 0 : full_rays, out_bytes, internal_rgb, dbg, npix_out);
    if (rc != RT_OK) return rc;
    const bool sm_affine = fused && (s->sm_affine || (flags & RT_FLAG_SM_AFFINE)) && s->sm_count <= RT_MAX_SM;
    if (sm_affine) {
        // chunk state of the persistent kernel, sized for the un-culled frame (same reason as above)
        const size_t nb = full_rays / 32u, nbx = fp.tile_w > 0 ? 0 : (size_t)(fp.w + 7) / 8, nby = fp.tile_w > 0 ? 0 : (size_t)(fp.h + 3) / 4;
        const size_t words = 2 + RT_MAX_SM + nb / RT_CHUNK_BLOCKS + nbx + nby + 8;
        if (words > L.cap_chunk_state) {
            cudaFree(L.d_chunk_state); L.d_chunk_state = nullptr; L.cap_chunk_state = 0;
            CU(s, cudaMalloc((void **)&L.d_chunk_state, words * sizeof(unsigned int)));
            L.cap_chunk_state = words;
        }
    }
    cudaStream_t st = out.ext_stream ? out.ext_stream : L.stream;
    uint8_t *d_rgb = internal_rgb ? L.d_rgb : out.d_rgb;

    RenderJob job;
    memset(&job, 0, sizeof job);
    job.fp = fp;
    job.d_meshes = s->d_meshes; job.d_objs = L.d_objs; job.d_lights = L.d_lights;
    job.fr = L.inline_frame;
    job.precision = s->precision; job.flags = flags & 0x00ffffffu;
    job.d_hits = L.d_hits; job.d_hit_colour = L.d_hit_colour; job.d_hit_lit = L.d_hit_lit;
    job.d_hit_geom = L.d_hit_geom; job.split_shade = s->split_shade && !s->shade32; job.shade32 = s->shade32;
    job.d_counters = L.d_counters; job.d_stats = L.d_stats;
    job.d_out_rgb = d_rgb;
    job.d_out_colour = out.h_colour ? L.d_dbg_colour : nullptr;
    job.d_out_face = out.h_face ? L.d_dbg_face : nullptr;
    job.d_out_obj = out.h_obj ? L.d_dbg_obj : nullptr;
    job.d_out_lit = out.h_lit ? L.d_dbg_lit : nullptr;
    job.sm_count = s->sm_count;
    job.single_mesh = L.h_objs.size() == 1;
    if (job.single_mesh) {
        job.h_mesh0 = s->h_views[L.h_objs[0].mesh];
        for (int k = 0; k < 3; k++) job.h_obj0_pos[k] = L.h_objs[0].pos[k];
    }
    job.shade_minb = s->shade_minb;
    job.primary_persistent = s->primary_persistent;
    job.fused = fused;
    job.d_tune = s->d_tune; job.d_tune_host = s->d_tune_host;
    job.any_plain = (s->h_tune ? *(volatile unsigned int *)s->h_tune : (unsigned int)s->any_defer_fixed) == 1u;
    job.sm_affine = sm_affine; job.d_chunk_state = L.d_chunk_state; job.chunk_state_words = L.cap_chunk_state;

    if (out.first) {
        CU(s, cudaMemsetAsync(L.d_counters, 0, 4 * sizeof(unsigned int), st));
        if (flags & RT_FLAG_COUNT) CU(s, cudaMemsetAsync(L.d_stats, 0, 2 * sizeof(Counters), st));
    }
    // background: sequential/main.go:17 surface.FillRect(nil, 0); distributed/main.go:70 default 0,0,0
    if (fp.tile_final) CU(s, rt_launch_clear_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, d_rgb, st));   // only this rank's pixels (peer framebuffer)
    else CU(s, cudaMemsetAsync(d_rgb, 0, out_bytes, st));
    if (job.d_out_colour) CU(s, cudaMemsetAsync(job.d_out_colour, 0, sizeof(double) * 3 * npix_out, st));
    if (job.d_out_face) CU(s, cudaMemsetAsync(job.d_out_face, 0xff, sizeof(int32_t) * npix_out, st));
    if (job.d_out_obj) CU(s, cudaMemsetAsync(job.d_out_obj, 0xff, sizeof(int32_t) * npix_out, st));
    if (job.d_out_lit) CU(s, cudaMemsetAsync(job.d_out_lit, 0, sizeof(uint32_t) * npix_out, st));

    auto cancelled = [&]() { return s->cancel_gen.load() != my_gen; };
    if (out.first) CU(s, cudaEventRecord(L.ev[0], st));
    if (job.fused) {
        CU(s, rt_launch_fused(job, st));
        if (out.last) { CU(s, cudaEventRecord(L.ev[1], st)); CU(s, cudaEventRecord(L.ev[2], st)); }
    } else {
        CU(s, rt_launch_primary(job, st));
        CU(s, cudaEventRecord(L.ev[1], st));
        if (cancelled()) { cudaStreamSynchronize(st); return fail(s, RT_ERR_CANCELLED, "cancelled"); }
        CU(s, rt_launch_shade(job, st));
        CU(s, cudaEventRecord(L.ev[2], st));
        if (cancelled()) { cudaStreamSynchronize(st); return fail(s, RT_ERR_CANCELLED, "cancelled"); }
        CU(s, rt_launch_write(job, st));
    }
    if (out.last) CU(s, cudaEventRecord(L.ev[3], st));

    L.stats.primary_rays = (out.first ? 0 : L.stats.primary_rays) + (uint64_t)fp.w * (uint64_t)fp.h;
    if (fp.tile_w > 0) L.stats.primary_rays = (uint64_t)fp.n_local_tiles * (uint64_t)(fp.tile_w * fp.tile_h);   // answered pixels, traced or culled
    L.stats_pending = true;
    L.stats_counted = (flags & RT_FLAG_COUNT) != 0;
    L.last_stream = st;
    if (out.async) return RT_OK;

    // copy results to the caller's host buffers
    if (out.h_rgb) {
        cudaPointerAttributes attr;
        bool pinned = cudaPointerGetAttributes(&attr, out.h_rgb) == cudaSuccess &&
                      (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        if (pinned) {
            CU(s, cudaMemcpyAsync(out.h_rgb, d_rgb, out_bytes, cudaMemcpyDeviceToHost, st));
            CU(s, cudaStreamSynchronize(st));
        } else {
            if (out_bytes > L.cap_pinned) {
                if (L.h_pinned) cudaFreeHost(L.h_pinned);
                L.h_pinned = nullptr; L.cap_pinned = 0;
                CU(s, cudaMallocHost((void **)&L.h_pinned, out_bytes));
                L.cap_pinned = out_bytes;
            }
            CU(s, cudaMemcpyAsync(L.h_pinned, d_rgb, out_bytes, cudaMemcpyDeviceToHost, st));
            CU(s, cudaStreamSynchronize(st));
            memcpy(out.h_rgb, L.h_pinned, out_bytes);
        }
    }
    if (out.h_colour) CU(s, cudaMemcpyAsync(out.h_colour, job.d_out_colour, sizeof(double) * 3 * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_face) CU(s, cudaMemcpyAsync(out.h_face, job.d_out_face, sizeof(int32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_obj) CU(s, cudaMemcpyAsync(out.h_obj, job.d_out_obj, sizeof(int32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_lit) CU(s, cudaMemcpyAsync(out.h_lit, job.d_out_lit, sizeof(uint32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    CU(s, cudaStreamSynchronize(st));
    // the kernels poll the cancel word and give up early: whatever is in the buffers then is not a frame
    if (cancelled()) return fail(s, RT_ERR_CANCELLED, "cancelled");
    // counters and event times are read lazily by rt_get_stats: two blocking device->host copies per call would be on
    // the critical path of concurrent BulkTrace callers
    return RT_OK;
}

// fills the culling fields of fp (rt_cull.h) from the lane's object positions
void cull_params(rt_scene *s, const Lane &L, FrameParams &fp, uint32_t flags) {
    if (!s->cull || (flags & (RT_FLAG_NO_CULL | RT_FLAG_BRUTE_FORCE))) { rt_cull_full(fp); return; }
    rt_cull_frame(fp, s->h_views.data(), L.h_objs.data(), (int)L.h_objs.size());
}

int region_params(rt_scene *s, const Lane &L, FrameParams &fp, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
                  uint32_t flags) {
    if (W == 0 || H == 0 || w == 0 || h == 0) return fail(s, RT_ERR_INVALID, "empty screen or region");
    if (W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "screen larger than 65536 pixels on a side");
    if ((uint64_t)x + w > W || (uint64_t)y + h > H) return fail(s, RT_ERR_INVALID, "region exceeds the screen");
    fp = L.frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = (int)x; fp.y = (int)y; fp.w = (int)w; fp.h = (int)h;
    fp.tile_w = fp.tile_h = fp.tiles_x = fp.rank = fp.n_ranks = fp.n_local_tiles = fp.tile_final = 0;
    cull_params(s, L, fp, flags);
    return RT_OK;
}

// tile mode over the region (x, y, w, h) of the W x H screen: the tiles are cut from the region
int tile_params_region(rt_scene *s, const Lane &L, FrameParams &fp, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
                       uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks, bool need_smem_fit, uint32_t flags) {
    if (W == 0 || H == 0 || W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "bad screen size");
    if (w == 0 || h == 0 || (uint64_t)x + w > W || (uint64_t)y + h > H) return fail(s, RT_ERR_INVALID, "empty region or region exceeds the screen");
    if (tile_w == 0 || tile_h == 0 || tile_w % 8 || tile_h % 4) return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4");
    if (need_smem_fit && (size_t)tile_w * tile_h * 3 > 48 * 1024) return fail(s, RT_ERR_INVALID, "tile larger than 16384 pixels");
    if (n_ranks == 0 || rank >= n_ranks) return fail(s, RT_ERR_INVALID, "bad rank / n_ranks");
    fp = L.frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = (int)x; fp.y = (int)y; fp.w = (int)w; fp.h = (int)h;
    fp.tile_w = (int)tile_w; fp.tile_h = (int)tile_h;
    fp.tiles_x = (int)((w + tile_w - 1) / tile_w);
    const uint32_t n_tiles = (uint32_t)fp.tiles_x * ((h + tile_h - 1) / tile_h);
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)((n_tiles > rank) ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    fp.tile_final = 0;
    cull_params(s, L, fp, flags);
    return RT_OK;
}
int tile_params(rt_scene *s, const Lane &L, FrameParams &fp, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank,
                uint32_t n_ranks, bool need_smem_fit, uint32_t flags) {
    return tile_params_region(s, L, fp, W, H, 0, 0, W, H, tile_w, tile_h, rank, n_ranks, need_smem_fit, flags);
}

// the device API on a lane other than 0 inherits lane 0's frame state until it is given its own (rt_frame_set_lane)
void inherit_frame(rt_scene *s, Lane &L, int idx) {
    if (idx == 0 || L.have_frame) return;
    Lane &L0 = *s->lanes[0];
    if (!L0.have_frame || !L0.inline_frame.use_lights || !L0.inline_frame.use_objs) return;
    L.frame = L0.frame; L.inline_frame = L0.inline_frame; L.h_objs = L0.h_objs; L.have_frame = true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int rt_internal_fail(rt_scene *s, int code, const std::string &msg) { return fail(s, code, msg); }

// The stripes of rank `rank` of the REGION (x, y, w, h): rendered in place into d_local_framebuffer (w*h*3, the region's final layout), then
// moved with ONE strided copy-engine transfer into h_framebuffer (the same layout in host memory).  rt_render_stripes_host is the
// full-screen case; rt_group_bulk_trace (rt_group.cu) drives one such call per GPU of the box for one BulkTrace order.
int rt_internal_render_stripes(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                               uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t stripe, uint32_t rank, uint32_t n_ranks,
                               void *d_local_framebuffer, void *h_framebuffer, void *d_sync_slot, uint32_t seq, void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_local_framebuffer || !h_framebuffer || !stream) return fail(s, RT_ERR_INVALID, "rt_render_stripes_host: NULL buffer or stream");
    if (n_ranks > 32 && d_sync_slot) return fail(s, RT_ERR_INVALID, "rt_render_stripes_host: the hand-off record holds 32 ranks");
    const bool rowmajor = (flags & RT_LAYOUT_ROWMAJOR) != 0;
    if (stripe == 0 || stripe % (rowmajor ? 4u : 8u)) return fail(s, RT_ERR_INVALID, "stripe must be a multiple of 8 columns (4 rows)");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    int rc = frame_set_lane(s, *Lp, cam, lights, n_lights, obj_pos);
    if (rc != RT_OK) return rc;
    const uint32_t tile_w = rowmajor ? ((w + 7u) / 8u) * 8u : stripe, tile_h = rowmajor ? stripe : ((h + 3u) / 4u) * 4u;
    FrameParams fp;
    rc = tile_params_region(s, *Lp, fp, W, H, x, y, w, h, tile_w, tile_h, rank, n_ranks, false, flags);
    if (rc != RT_OK) return rc;
    fp.tile_final = 1;
    cudaStream_t st = (cudaStream_t)stream;
    RenderOut o; o.d_rgb = (uint8_t *)d_local_framebuffer; o.ext_stream = st; o.async = true;
    s->last_lane.store(idx);
    rc = render_lane(s, *Lp, fp, flags, (size_t)w * h, o);
    if (rc != RT_OK) return rc;
    if (d_sync_slot) CU(s, rt_launch_wait_consumed((uint32_t *)d_sync_slot, seq, rank, s->sync_timeout_ms, st));
    // stripe t = rank + k*n_ranks starts at byte t*run; all but a cut last stripe are `run` bytes long
    const size_t extent = rowmajor ? h : w, other = rowmajor ? w : h;
    const size_t run = (size_t)stripe * other * 3;
    const size_t n_stripes = (extent + stripe - 1) / stripe, n_full = extent / stripe;
    const size_t mine_full = n_full > rank ? (n_full - rank + n_ranks - 1) / n_ranks : 0;
    const uint8_t *src = (const uint8_t *)d_local_framebuffer;
    uint8_t *dst = (uint8_t *)h_framebuffer;
    if (mine_full)
        CU(s, cudaMemcpy2DAsync(dst + (size_t)rank * run, (size_t)n_ranks * run, src + (size_t)rank * run, (size_t)n_ranks * run, run, mine_full,
                                cudaMemcpyDeviceToHost, st));
    if (n_stripes > n_full && (n_stripes - 1) % n_ranks == rank) {
        const size_t off = (n_stripes - 1) * run, bytes = (extent - n_full * stripe) * other * 3;
        CU(s, cudaMemcpyAsync(dst + off, src + off, bytes, cudaMemcpyDeviceToHost, st));
    }
    if (d_sync_slot) CU(s, rt_launch_signal_arrive((uint32_t *)d_sync_slot, seq, rank, st));
    return RT_OK;
}


extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *rt_last_error(const rt_scene *scene) {
    if (!scene) return g_err.c_str();
    if (!g_err.empty() && g_err_scene == scene) return g_err.c_str();       // this thread's last call failed on this handle
    { std::lock_guard<std::mutex> lk(const_cast<rt_scene *>(scene)->mu); g_err_ret = scene->err; }
    return g_err_ret.c_str();
}

// Common part of the two registration entry points.  load_mesh(k, m, st, err) must leave mesh k's per-face records on the device
// (m.fverts, m.fnormals or nullptr, m.fmat, m.materials, m.n_faces, m.n_materials); the LBVHs are built here.
static int create_scene_common(uint32_t precision, int32_t device, uint32_t n_meshes, const double *obj_pos, const uint32_t *obj_mesh,
                               uint32_t n_objects, const std::function<int(uint32_t, MeshDevice &, cudaStream_t, std::string &)> &load_mesh,
                               rt_scene **out) {
    *out = nullptr;
    if (precision != RT_FP64_EXACT && precision != RT_FP32_FAST) return fail(nullptr, RT_ERR_INVALID, "unknown precision mode");
    if (n_objects && (!obj_pos || !obj_mesh)) return fail(nullptr, RT_ERR_INVALID, "objects are NULL");
    for (uint32_t k = 0; k < n_objects; k++)
        if (obj_mesh[k] >= n_meshes) return fail(nullptr, RT_ERR_INVALID, "object refers to a mesh out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, RT_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, RT_ERR_INVALID, "device ordinal out of range");
    rt_scene *s = new (std::nothrow) rt_scene();
    if (!s) return fail(nullptr, RT_ERR_OOM, "out of host memory");
    s->device = device;
    s->precision = precision;
    s->max_leaf = env_int("RT_B200_MAX_LEAF", 4);
    s->shade_minb = env_int("RT_B200_SHADE_MINB", 9);   // fused kernel: 9 CTAs/SM = 56 registers (measured best); K4 of the pipeline caps at 8
    s->primary_persistent = env_int("RT_B200_PERSISTENT", 0) != 0;
    s->split_shade = env_int("RT_B200_SPLIT_SHADE", 0) != 0;
    s->shade32 = env_int("RT_B200_SHADE32", 0) != 0 && precision == RT_FP32_FAST;
    s->max_leaf_any = env_int("RT_B200_MAX_LEAF_SHADE", RT_DEFAULT_MAX_LEAF_SHADE);
    s->fused = env_int("RT_B200_FUSED", RT_DEFAULT_FUSED) != 0;
    s->sm_affine = env_int("RT_B200_SM_AFFINE", 0) != 0;
    s->bulk_strips = env_int("RT_B200_BULK_STRIPS", 4);
    s->cull = env_int("RT_B200_CULL", 1) != 0;
    s->sync_timeout_ms = (uint32_t)std::max(1, env_int("RT_B200_SYNC_TIMEOUT_MS", 2000));
    int n_lanes = env_int("RT_B200_LANES", 9);      // lane 0 + a pool of 8 for concurrent BulkTrace calls / frames in flight
    if (n_lanes < 1) n_lanes = 1;
    if (n_lanes > 64) n_lanes = 64;
    auto bail = [&](int code, const std::string &msg) { destroy_all(s); delete s; g_err = msg; return code; };
#define CUB_(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    CUB_(cudaSetDevice(s->device));
    {
        cudaDeviceProp prop;
        CUB_(cudaGetDeviceProperties(&prop, s->device));
        s->sm_count = prop.multiProcessorCount;
        // the kernels are compiled for sm_100a only: on any other device the first launch fails with
        // cudaErrorNoKernelImageForDevice and the error is returned (no fallback)
    }
    s->h_objs.resize(n_objects);
    for (uint32_t k = 0; k < n_objects; k++) {
        for (int c = 0; c < 3; c++) s->h_objs[k].pos[c] = obj_pos[3 * (size_t)k + c];
        s->h_objs[k].mesh = obj_mesh[k]; s->h_objs[k].pad = 0;
    }
    for (int k = 0; k < n_lanes; k++) {
        s->lanes.emplace_back(new Lane());
        Lane &L = *s->lanes.back();
        memset(&L.stats, 0, sizeof L.stats);
        memset(&L.frame, 0, sizeof L.frame);
        memset(&L.inline_frame, 0, sizeof L.inline_frame);
        L.h_objs = s->h_objs;
        CUB_(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        for (auto &e : L.ev) CUB_(cudaEventCreate(&e));
    }
    CUB_(cudaStreamCreateWithFlags(&s->cancel_stream, cudaStreamNonBlocking));
    CUB_(cudaMallocHost((void **)&s->h_cancel_word, sizeof(unsigned int)));
    if (!getenv("RT_B200_ANY_DEFER") && s->precision == RT_FP32_FAST) {       // self-tuned unless the caller fixes the bound
        CUB_(cudaMalloc((void **)&s->d_tune, 4 * sizeof(unsigned int)));
        const unsigned int init[4] = { 0u, 0u, 8u, 0u };
        CUB_(cudaMemcpy(s->d_tune, init, sizeof init, cudaMemcpyHostToDevice));
        CUB_(cudaHostAlloc((void **)&s->h_tune, sizeof(unsigned int), cudaHostAllocMapped));
        *s->h_tune = 8u;
        CUB_(cudaHostGetDevicePointer((void **)&s->d_tune_host, s->h_tune, 0));
    } else s->any_defer_fixed = std::max(1, std::min(env_int("RT_B200_ANY_DEFER", 8), RT_STACK_SIZE));
    cudaStream_t st0 = s->lanes[0]->stream;
    s->meshes.resize(n_meshes);
    s->h_views.resize(n_meshes);
    for (uint32_t k = 0; k < n_meshes; k++) {
        MeshDevice &m = s->meshes[k];
        std::string e;
        int rc = load_mesh(k, m, st0, e);
        if (rc != RT_OK) return bail(rc, e);
        CUB_(cudaStreamSynchronize(st0));
        // the any-hit hierarchy is only walked by FP32_FAST shadow rays
        const rt_bvh_info ingest = m.info;           // rt_build_bvh resets the record
        if (!rt_build_bvh(m, s->precision == RT_FP32_FAST, s->precision == RT_FP64_EXACT, s->max_leaf,
                          s->precision == RT_FP32_FAST ? s->max_leaf_any : s->max_leaf, s->sm_count, st0, e))
            return bail(RT_ERR_CUDA, "BVH build: " + e);
        m.info.ms_ingest = ingest.ms_ingest; m.info.n_verts_ingested = ingest.n_verts_ingested; m.info.n_normals_ingested = ingest.n_normals_ingested;
        MeshView &v = s->h_views[k];
        v.nodes = m.nodes; v.tri32 = m.tri32; v.tri64 = m.tri64; v.fverts = m.fverts; v.fnormals = m.fnormals; v.fmat = m.fmat;
        v.materials = m.materials; v.n_faces = m.n_faces; v.n_nodes = m.n_nodes; v.root_ref = m.root_ref;
        v.nodes_any = m.nodes_any; v.root_ref_any = m.root_ref_any; v.n_nodes_any = m.n_nodes_any;
        v.qnodes = m.qnodes; v.qnodes_any = m.qnodes_any;
        v.any_defer = (uint32_t)std::max(1, std::min(env_int("RT_B200_ANY_DEFER", 8), RT_STACK_SIZE));
        for (int c = 0; c < 3; c++) { v.lo[c] = m.lo[c]; v.hi[c] = m.hi[c]; v.vlo[c] = m.vlo[c]; v.vhi[c] = m.vhi[c]; v.qorg[c] = m.qorg[c]; v.qstep[c] = m.qstep[c]; }
    }
    CUB_(upload(&s->d_meshes, s->h_views.data(), s->h_views.size(), st0));
    CUB_(cudaStreamSynchronize(st0));
#undef CUB_
    *out = s;
    return RT_OK;
}

static const char *bad_index_message(int bad) {
    return bad == 1 ? "mesh: vertex index out of range" : bad == 2 ? "mesh: normal index out of range" : "mesh: material index out of range";
}

int rt_scene_create(const rt_scene_desc *desc, rt_scene **out) {
    API_ENTER();
    if (!desc || !out) return fail(nullptr, RT_ERR_INVALID, "rt_scene_create: NULL argument");
    *out = nullptr;
    if (desc->n_meshes && !desc->meshes) return fail(nullptr, RT_ERR_INVALID, "meshes is NULL");
    auto load = [&](uint32_t k, MeshDevice &m, cudaStream_t st0, std::string &e) -> int {
        const rt_mesh_desc &d = desc->meshes[k];
        int rc = check_mesh_desc(d, e);
        if (rc != RT_OK) return rc;
        m.n_faces = d.n_faces; m.n_materials = d.n_materials;
        // the indexed mesh goes up as it is (C3: 52 MB instead of 147 MB of per-face records) and is gathered on the device
        double *d_verts = nullptr, *d_normals = nullptr; uint32_t *d_faces = nullptr;
        cudaError_t ue = upload(&d_verts, d.verts, (size_t)3 * d.n_verts, st0);
        if (ue == cudaSuccess && d.n_normals) ue = upload(&d_normals, d.normals, (size_t)3 * d.n_normals, st0);
        if (ue == cudaSuccess) ue = upload(&d_faces, d.faces, (size_t)7 * d.n_faces, st0);
        int bad = 0;
        if (ue == cudaSuccess) ue = rt_flatten_faces(m, d_verts, d.n_verts, d_normals, d.n_normals, d_faces, st0, &bad);
        cudaFree(d_verts); cudaFree(d_normals); cudaFree(d_faces);
        if (ue == cudaSuccess) ue = upload(&m.materials, d.materials, (size_t)10 * d.n_materials, st0);
        if (ue != cudaSuccess) { e = std::string("scene upload: ") + cudaGetErrorString(ue); return ue == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA; }
        if (bad) { e = bad_index_message(bad); return RT_ERR_INVALID; }
        return RT_OK;
    };
    return create_scene_common(desc->precision, desc->device, desc->n_meshes, desc->obj_pos, desc->obj_mesh, desc->n_objects, load, out);
}

// registration straight from the OBJ text.  The file's bytes are parsed ON THE GPU (obj_ingest.cu) into the
// indexed arrays rt_mesh_desc describes -- no gwob tokenizer, no value-keyed de-duplication maps (mesh.go:146-203), no R-tree
// Insert per face (mesh.go:208) -- and flow on the device into the same flatten + LBVH build as rt_scene_create.
int rt_scene_create_obj(const rt_obj_mesh_desc *meshes, uint32_t n_meshes, const double *obj_pos, const uint32_t *obj_mesh, uint32_t n_objects,
                        int32_t device, uint32_t precision, rt_scene **out) {
    API_ENTER();
    if (!out || (n_meshes && !meshes)) return fail(nullptr, RT_ERR_INVALID, "rt_scene_create_obj: NULL argument");
    *out = nullptr;
    auto load = [&](uint32_t k, MeshDevice &m, cudaStream_t st0, std::string &e) -> int {
        const rt_obj_mesh_desc &d = meshes[k];
        if (d.obj_bytes && !d.obj_text) { e = "mesh: obj_text is NULL"; return RT_ERR_INVALID; }
        if (d.n_materials == 0 || !d.materials) { e = "mesh: no materials"; return RT_ERR_INVALID; }
        if (d.n_usemtl && (!d.usemtl_offset || !d.usemtl_material)) { e = "mesh: usemtl table is NULL"; return RT_ERR_INVALID; }
        if (d.default_material >= d.n_materials) { e = "mesh: material index out of range"; return RT_ERR_INVALID; }
        for (uint32_t u = 0; u < d.n_usemtl; u++)
            if (d.usemtl_material[u] >= d.n_materials || (u && d.usemtl_offset[u] < d.usemtl_offset[u - 1])) { e = "mesh: bad usemtl table"; return RT_ERR_INVALID; }
        double *d_verts = nullptr, *d_normals = nullptr; uint32_t *d_faces = nullptr;
        uint32_t nv = 0, nn = 0, nf = 0;
        float ms = 0.f;
        if (!rt_obj_ingest(d.obj_text, d.obj_bytes, d.usemtl_offset, d.usemtl_material, d.n_usemtl, d.default_material, st0, &d_verts, &nv, &d_normals,
                           &nn, &d_faces, &nf, &ms, e)) {
            e = "OBJ ingest: " + e;
            return RT_ERR_INVALID;
        }
        m.n_faces = nf; m.n_materials = d.n_materials;
        int bad = 0;
        cudaError_t ue = rt_flatten_faces(m, d_verts, nv, nn ? d_normals : nullptr, nn, d_faces, st0, &bad);
        cudaFree(d_verts); cudaFree(d_normals); cudaFree(d_faces);
        if (ue == cudaSuccess) ue = upload(&m.materials, d.materials, (size_t)10 * d.n_materials, st0);
        if (ue != cudaSuccess) { e = std::string("scene upload: ") + cudaGetErrorString(ue); return ue == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA; }
        if (bad) { e = bad_index_message(bad); return RT_ERR_INVALID; }
        m.info.ms_ingest = ms;
        m.info.n_verts_ingested = nv; m.info.n_normals_ingested = nn;
        return RT_OK;
    };
    return create_scene_common(precision, device, n_meshes, obj_pos, obj_mesh, n_objects, load, out);
}

void rt_scene_destroy(rt_scene *scene) {
    if (!scene) return;
    for (auto &lp : scene->lanes) { std::lock_guard<std::mutex> lk(lp->mu); }    // wait for calls in flight on every lane
    destroy_all(scene);
    delete scene;
}

int rt_lane_count(rt_scene *s) { return s ? (int)s->lanes.size() : 0; }

int rt_frame_set(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos) {
    API_ENTER();
    return rt_frame_set_lane(s, 0, cam, lights, n_lights, obj_pos);
}

int rt_frame_set_lane(rt_scene *s, uint32_t lane, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (lane >= s->lanes.size()) return fail(s, RT_ERR_INVALID, "lane index out of range");
    Lane &L = *s->lanes[lane];
    std::lock_guard<std::mutex> lk(L.mu);
    return frame_set_lane(s, L, cam, lights, n_lights, obj_pos);
}

int rt_render_region_debug(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                           double *out_colour, int32_t *out_face, int32_t *out_obj, uint32_t *out_lit, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    Lane &L = *s->lanes[0];
    std::lock_guard<std::mutex> lk(L.mu);
    FrameParams fp;
    int rc = region_params(s, L, fp, W, H, x, y, w, h, flags);
    if (rc != RT_OK) return rc;
    RenderOut o; o.h_rgb = out_rgb; o.h_colour = out_colour; o.h_face = out_face; o.h_obj = out_obj; o.h_lit = out_lit;
    s->last_lane.store(0);
    return render_lane(s, L, fp, flags, (size_t)w * h, o);
}

int rt_render_region(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                     uint32_t flags) {
    API_ENTER();
    if (!out_rgb) return fail(s, RT_ERR_INVALID, "out_rgb is NULL");
    return rt_render_region_debug(s, W, H, x, y, w, h, out_rgb, nullptr, nullptr, nullptr, nullptr, flags);
}

// One BulkTrace order answered in STRIPS (RT_B200_BULK_STRIPS, default 4; orders below 256k pixels are not cut): the reply of a 4K order
// is 24.9 MB, whose device->host copy (~0.45 ms) used to start when the whole frame kernel (~0.74 ms) had finished.  The region is
// cut into strips that are contiguous in the reply (BulkTrace's column-major layout: ranges of pixel columns, a multiple of 8 wide;
// RT_LAYOUT_ROWMAJOR: ranges of rows, a multiple of 4 high); the strips render together on streams of descending priority, and strip k crosses PCIe on
// the lane's copy stream as soon as it is done, while the later strips are still rendering.  Same kernels, same 8x4 blocks relative to the strip's origin, same bytes.
static int bulk_strips(rt_scene *s, uint32_t w, uint32_t h, uint32_t flags) {
    // only frames of the fused kernel are cut: its one counter (hit pixels) simply keeps adding up across the strips, whereas the hit list of
    // the three-kernel pipeline is per launch; FP64_EXACT orders are compute-bound by orders of magnitude
    if (s->bulk_strips <= 1 || (uint64_t)w * h < (1u << 18) || !will_fuse(s, flags)) return 1;
    const bool rowmajor = (flags & RT_LAYOUT_ROWMAJOR) != 0;
    const uint32_t unit = rowmajor ? 4u : 8u, extent = rowmajor ? h : w;
    const uint32_t most = extent / (8u * unit);                    // a strip is at least 8 blocks thick
    return (int)std::min<uint32_t>((uint32_t)s->bulk_strips, most ? most : 1u);
}

static int bulk_trace_strips(rt_scene *s, Lane &L, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                             uint32_t flags, int strips) {
    CU(s, cudaSetDevice(s->device));
    const size_t total = (size_t)w * h * 3;
    int rc = ensure_scratch(s, L, 0, total, true, false, (size_t)w * h);
    if (rc != RT_OK) return rc;
    if (!L.copy_stream) {
        CU(s, cudaStreamCreateWithFlags(&L.copy_stream, cudaStreamNonBlocking));
        CU(s, cudaEventCreateWithFlags(&L.start_ev, cudaEventDisableTiming));
    }
    int prio_least = 0, prio_greatest = 0;
    CU(s, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));          // greatest = numerically lowest
    while ((int)L.strip_ev.size() < strips) {
        const int k = (int)L.strip_ev.size();
        cudaEvent_t a = nullptr, b = nullptr;
        cudaStream_t t = nullptr;
        CU(s, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        L.strip_ev.push_back(a);
        CU(s, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        L.copy_ev.push_back(b);
        CU(s, cudaStreamCreateWithPriority(&t, cudaStreamNonBlocking, std::min(prio_greatest + k, prio_least)));
        L.strip_stream.push_back(t);
    }
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, out_rgb) == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (!pinned && total > L.cap_pinned) {
        if (L.h_pinned) cudaFreeHost(L.h_pinned);
        L.h_pinned = nullptr; L.cap_pinned = 0;
        CU(s, cudaMallocHost((void **)&L.h_pinned, total));
        L.cap_pinned = total;
    }
    uint8_t *dst = pinned ? out_rgb : L.h_pinned;
    const bool rowmajor = (flags & RT_LAYOUT_ROWMAJOR) != 0;
    const uint32_t unit = rowmajor ? 4u : 8u, extent = rowmajor ? h : w;
    const uint32_t units = (extent + unit - 1) / unit;
    const unsigned int my_gen = s->cancel_gen.load();
    // the frame's counters and its first timing event belong to the lane's own stream; every strip starts after them
    CU(s, cudaMemsetAsync(L.d_counters, 0, 4 * sizeof(unsigned int), L.stream));
    if (flags & RT_FLAG_COUNT) CU(s, cudaMemsetAsync(L.d_stats, 0, 2 * sizeof(Counters), L.stream));
    CU(s, cudaEventRecord(L.ev[0], L.stream));
    CU(s, cudaEventRecord(L.start_ev, L.stream));
    L.stats.primary_rays = 0;
    std::vector<size_t> off(strips + 1, 0);
    auto drain = [&]() { for (int k = 0; k < strips; k++) cudaStreamSynchronize(L.strip_stream[k]); cudaStreamSynchronize(L.copy_stream); cudaStreamSynchronize(L.stream); };
    for (int k = 0; k < strips; k++) {
        const uint32_t a = (uint32_t)((uint64_t)units * k / strips) * unit, b = k + 1 == strips ? extent : (uint32_t)((uint64_t)units * (k + 1) / strips) * unit;
        FrameParams fp;
        rc = rowmajor ? region_params(s, L, fp, W, H, x, y + a, w, b - a, flags) : region_params(s, L, fp, W, H, x + a, y, b - a, h, flags);
        if (rc != RT_OK) { drain(); return rc; }
        off[k] = (size_t)a * (rowmajor ? w : h) * 3;
        off[k + 1] = (size_t)b * (rowmajor ? w : h) * 3;
        // the strips are in flight TOGETHER (the GPU stays full: one strip alone is mostly kernel tail), each on its own stream; the
        // block scheduler serves the streams in priority order, so strip k finishes -- and starts crossing PCIe -- before strip k+1
        cudaStream_t st = L.strip_stream[k];
        CU(s, cudaStreamWaitEvent(st, L.start_ev, 0));
        RenderOut o;
        o.d_rgb = L.d_rgb + off[k]; o.async = true; o.first = false; o.last = false; o.ext_stream = st;
        rc = render_lane(s, L, fp, flags, (size_t)(b - a) * (rowmajor ? w : h), o);
        if (rc != RT_OK) { drain(); return rc; }
        CU(s, cudaEventRecord(L.strip_ev[k], st));
        CU(s, cudaStreamWaitEvent(L.copy_stream, L.strip_ev[k], 0));
        CU(s, cudaMemcpyAsync(dst + off[k], L.d_rgb + off[k], off[k + 1] - off[k], cudaMemcpyDeviceToHost, L.copy_stream));
        if (!pinned) CU(s, cudaEventRecord(L.copy_ev[k], L.copy_stream));
        CU(s, cudaStreamWaitEvent(L.stream, L.strip_ev[k], 0));
    }
    CU(s, cudaEventRecord(L.ev[1], L.stream)); CU(s, cudaEventRecord(L.ev[2], L.stream)); CU(s, cudaEventRecord(L.ev[3], L.stream));
    L.last_stream = L.stream;
    if (pinned) CU(s, cudaStreamSynchronize(L.copy_stream));
    else for (int k = 0; k < strips; k++) {                       // pageable destination: strip k is copied out while the later ones are in flight
        CU(s, cudaEventSynchronize(L.copy_ev[k]));
        memcpy(out_rgb + off[k], L.h_pinned + off[k], off[k + 1] - off[k]);
    }
    CU(s, cudaStreamSynchronize(L.stream));
    if (s->cancel_gen.load() != my_gen) return fail(s, RT_ERR_CANCELLED, "cancelled");
    return RT_OK;
}

int rt_bulk_trace(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                  uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!out_rgb) return fail(s, RT_ERR_INVALID, "out_rgb is NULL");
    int idx = 0;
    Lane *Lp = acquire_lane(s, &idx);          // a free pool lane: concurrent calls run concurrently on the GPU
    int rc;
    {
        std::lock_guard<std::mutex> lk(Lp->mu);
        // an order without object positions renders the registered ones (LinkTo, environment.go:73-98)
        if (!obj_pos) Lp->h_objs = s->h_objs;
        rc = frame_set_lane(s, *Lp, cam, lights, n_lights, obj_pos);
        FrameParams fp;
        if (rc == RT_OK) rc = region_params(s, *Lp, fp, W, H, x, y, w, h, flags);
        if (rc == RT_OK) {
            s->last_lane.store(idx);
            // Strips buy latency for an order that has the GPU to itself (one order per frame and worker is how master/pool drives a
            // worker: 1.21 -> 1.00 ms for the 4K frame); with other orders in flight on this handle the GPU is full anyway, their copies
            // already overlap this order's kernel, and the extra launches cost 2.6 % of throughput -- those orders stay uncut.
            int others = 0;
            { std::lock_guard<std::mutex> lk2(s->mu); for (auto &lp : s->lanes) others += (lp->busy && lp.get() != Lp) ? 1 : 0; }
            const int strips = others ? 1 : bulk_strips(s, w, h, flags);
            if (strips > 1) rc = bulk_trace_strips(s, *Lp, W, H, x, y, w, h, out_rgb, flags, strips);
            else {
                RenderOut o; o.h_rgb = out_rgb;
                rc = render_lane(s, *Lp, fp, flags, (size_t)w * h, o);
            }
        }
    }
    release_lane(s, Lp);
    return rc;
}

// ---- result wire format: comms.TraceResults as proto.Marshal would emit it --------------------
uint64_t rt_trace_results_max_bytes(uint64_t n_pixels) { return 11ull * n_pixels; }     // 0x0A len + 3 x (tag + 2-byte varint)

static int ensure_wire_scratch(rt_scene *s, Lane &L, uint64_t n_pixels, bool need_out) {
    const size_t need = rt_wire_scratch_bytes(n_pixels);
    if (need > L.cap_wire_scratch) {
        cudaFree(L.d_wire_scratch); L.d_wire_scratch = nullptr; L.cap_wire_scratch = 0;
        CU(s, cudaMalloc(&L.d_wire_scratch, need));
        L.cap_wire_scratch = need;
    }
    if (!L.d_wire_len) {
        CU(s, cudaMalloc((void **)&L.d_wire_len, sizeof(uint64_t)));
        CU(s, cudaMallocHost((void **)&L.h_wire_len, sizeof(uint64_t)));
    }
    const size_t out_need = (size_t)rt_trace_results_max_bytes(n_pixels);
    if (need_out && out_need > L.cap_wire) {
        cudaFree(L.d_wire); L.d_wire = nullptr; L.cap_wire = 0;
        CU(s, cudaMalloc((void **)&L.d_wire, out_need ? out_need : 16));
        L.cap_wire = out_need;
    }
    return RT_OK;
}

int rt_encode_trace_results_device(rt_scene *s, const void *d_rgb, uint64_t n_pixels, void *d_out, uint64_t capacity, uint64_t *d_len,
                                   void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_len || (n_pixels && (!d_rgb || !d_out))) return fail(s, RT_ERR_INVALID, "rt_encode_trace_results_device: NULL argument");
    if (((uintptr_t)d_rgb & 3u) != 0) return fail(s, RT_ERR_INVALID, "rt_encode_trace_results_device: d_rgb must be 4-byte aligned");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    CU(s, cudaSetDevice(s->device));
    int rc = ensure_wire_scratch(s, *Lp, n_pixels, false);
    if (rc != RT_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : Lp->stream;
    CU(s, rt_launch_wire_encode((const uint8_t *)d_rgb, n_pixels, (uint8_t *)d_out, capacity, d_len, Lp->d_wire_scratch, st));
    if (!stream) CU(s, cudaStreamSynchronize(st));
    return RT_OK;
}

// One BulkTrace RPC down to the bytes grpc puts on the wire: frame state + render + encode on the device, then ONE
// device->host copy of exactly the message (the length is read back first, 8 bytes, while the encoder still runs).
int rt_bulk_trace_encoded(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                          uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out, uint64_t capacity, uint64_t *out_len,
                          uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!out || !out_len) return fail(s, RT_ERR_INVALID, "rt_bulk_trace_encoded: NULL argument");
    if (flags & RT_LAYOUT_ROWMAJOR) return fail(s, RT_ERR_INVALID, "TraceResults is in BulkTrace order (i*height + j): RT_LAYOUT_ROWMAJOR does not apply");
    int idx = 0;
    Lane *Lp = acquire_lane(s, &idx);
    int rc;
    {
        std::lock_guard<std::mutex> lk(Lp->mu);
        Lane &L = *Lp;
        if (!obj_pos) L.h_objs = s->h_objs;
        rc = frame_set_lane(s, L, cam, lights, n_lights, obj_pos);
        FrameParams fp;
        if (rc == RT_OK) rc = region_params(s, L, fp, W, H, x, y, w, h, flags);
        const uint64_t npix = (uint64_t)w * h;
        if (rc == RT_OK) {
            RenderOut o; o.async = true;                       // into the lane's own RGB buffer, no copy out
            s->last_lane.store(idx);
            rc = render_lane(s, L, fp, flags, (size_t)npix, o);
        }
        if (rc == RT_OK) rc = ensure_wire_scratch(s, L, npix, true);
        auto body = [&]() -> int {
            cudaStream_t st = L.stream;
            CU(s, rt_launch_wire_encode(L.d_rgb, npix, L.d_wire, L.cap_wire, L.d_wire_len, L.d_wire_scratch, st));
            // the length is final after the scan kernel; reading it here only waits for the encoder as well (same stream)
            CU(s, cudaMemcpyAsync(L.h_wire_len, L.d_wire_len, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CU(s, cudaStreamSynchronize(st));
            *out_len = *L.h_wire_len;
            if (*out_len > capacity) return fail(s, RT_ERR_INVALID, "rt_bulk_trace_encoded: out buffer too small (see *out_len)");
            cudaPointerAttributes attr;
            const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess &&
                                (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
            cudaGetLastError();
            if (pinned) {
                CU(s, cudaMemcpyAsync(out, L.d_wire, (size_t)*out_len, cudaMemcpyDeviceToHost, st));
                CU(s, cudaStreamSynchronize(st));
            } else {
                if (*out_len > L.cap_pinned) {
                    if (L.h_pinned) cudaFreeHost(L.h_pinned);
                    L.h_pinned = nullptr; L.cap_pinned = 0;
                    CU(s, cudaMallocHost((void **)&L.h_pinned, (size_t)*out_len));
                    L.cap_pinned = (size_t)*out_len;
                }
                CU(s, cudaMemcpyAsync(L.h_pinned, L.d_wire, (size_t)*out_len, cudaMemcpyDeviceToHost, st));
                CU(s, cudaStreamSynchronize(st));
                memcpy(out, L.h_pinned, (size_t)*out_len);
            }
            return RT_OK;
        };
        if (rc == RT_OK) rc = body();
    }
    release_lane(s, Lp);
    return rc;
}

int rt_render_region_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void *d_out_rgb,
                            void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_out_rgb) return fail(s, RT_ERR_INVALID, "d_out_rgb is NULL");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    inherit_frame(s, *Lp, idx);
    FrameParams fp;
    int rc = region_params(s, *Lp, fp, W, H, x, y, w, h, flags);
    if (rc != RT_OK) return rc;
    RenderOut o; o.d_rgb = (uint8_t *)d_out_rgb; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    s->last_lane.store(idx);
    rc = render_lane(s, *Lp, fp, flags, (size_t)w * h, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(Lp->stream)); collect_stats(s, *Lp); }
    return rc;
}

int rt_render_tiles_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                           void *d_out_tiles, void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_out_tiles) return fail(s, RT_ERR_INVALID, "d_out_tiles is NULL");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    inherit_frame(s, *Lp, idx);
    FrameParams fp;
    int rc = tile_params(s, *Lp, fp, W, H, tile_w, tile_h, rank, n_ranks, false, flags);
    if (rc != RT_OK) return rc;
    RenderOut o; o.d_rgb = (uint8_t *)d_out_tiles; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    s->last_lane.store(idx);
    rc = render_lane(s, *Lp, fp, flags & ~RT_LAYOUT_ROWMAJOR, (size_t)fp.n_local_tiles * tile_w * tile_h, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(Lp->stream)); collect_stats(s, *Lp); }
    return rc;
}

int rt_render_tiles_peer(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                         void *d_framebuffer, void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_framebuffer) return fail(s, RT_ERR_INVALID, "d_framebuffer is NULL");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    inherit_frame(s, *Lp, idx);
    FrameParams fp;
    int rc = tile_params(s, *Lp, fp, W, H, tile_w, tile_h, rank, n_ranks, false, flags);
    if (rc != RT_OK) return rc;
    fp.tile_final = 1;
    RenderOut o; o.d_rgb = (uint8_t *)d_framebuffer; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    s->last_lane.store(idx);
    rc = render_lane(s, *Lp, fp, flags, (size_t)W * H, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(Lp->stream)); collect_stats(s, *Lp); }
    return rc;
}

int rt_push_tiles_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                         const void *d_tiles, void *d_framebuffer, void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_tiles || !d_framebuffer) return fail(s, RT_ERR_INVALID, "rt_push_tiles_device: bad argument");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    CU(s, cudaSetDevice(s->device));
    FrameParams fp;
    int rc = tile_params(s, *Lp, fp, W, H, tile_w, tile_h, rank, n_ranks, true, flags);
    if (rc != RT_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : Lp->stream;
    CU(s, rt_launch_push_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, (const uint8_t *)d_tiles, (uint8_t *)d_framebuffer, PushSync{}, s->sm_count, st));
    if (!stream) CU(s, cudaStreamSynchronize(st));
    return RT_OK;
}

// rt_frame_set + rt_render_tiles_device + rt_push_tiles_device in one call (one lock, one ctypes/cgo crossing per frame)
int rt_render_push_tiles(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                         uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks, void *d_tiles, void *d_framebuffer,
                         void *d_sync_slot, uint32_t seq, void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_tiles || !d_framebuffer || !stream) return fail(s, RT_ERR_INVALID, "rt_render_push_tiles: NULL buffer or stream");
    int idx = 0;
    Lane *Lp = device_lane(s, flags, &idx);
    if (!Lp) return RT_ERR_INVALID;
    std::lock_guard<std::mutex> lk(Lp->mu);
    int rc = frame_set_lane(s, *Lp, cam, lights, n_lights, obj_pos);
    if (rc != RT_OK) return rc;
    FrameParams fp;
    rc = tile_params(s, *Lp, fp, W, H, tile_w, tile_h, rank, n_ranks, true, flags);
    if (rc != RT_OK) return rc;
    RenderOut o; o.d_rgb = (uint8_t *)d_tiles; o.ext_stream = (cudaStream_t)stream; o.async = true;
    s->last_lane.store(idx);
    rc = render_lane(s, *Lp, fp, flags & ~RT_LAYOUT_ROWMAJOR, (size_t)fp.n_local_tiles * tile_w * tile_h, o);
    if (rc != RT_OK) return rc;
    PushSync sy{};
    if (d_sync_slot) {
        if (n_ranks > 32) return fail(s, RT_ERR_INVALID, "rt_render_push_tiles: the hand-off record holds 32 ranks");
        sy.slot = (uint32_t *)d_sync_slot; sy.blocks_done = Lp->d_counters + RT_CNT_PUSH_DONE; sy.seq = seq; sy.timeout_ms = s->sync_timeout_ms; sy.debug = (uint32_t)env_int("RT_B200_SYNC_DEBUG", 0);
    }
    CU(s, rt_launch_push_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, (const uint8_t *)d_tiles, (uint8_t *)d_framebuffer, sy, s->sm_count, (cudaStream_t)stream));
    return RT_OK;
}

// Multi-GPU end-to-end path: the frame goes to HOST memory over every GPU's own PCIe link.  The screen is cut into stripes
// that are contiguous in the output layout (column-major BulkTrace layout: `stripe` pixel columns x full height;
// row-major: `stripe` pixel rows x full width), dealt round-robin; a rank renders its stripes in place into its local
// device framebuffer (final layout) and ONE strided copy-engine transfer moves them into the host framebuffer (pinned
// shared memory, rt_host_register); the hand-off flags live in the same host memory.
int rt_render_stripes_host(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                           uint32_t H, uint32_t stripe, uint32_t rank, uint32_t n_ranks, void *d_local_framebuffer, void *h_framebuffer,
                           void *d_sync_slot, uint32_t seq, void *stream, uint32_t flags) {
    API_ENTER();
    return rt_internal_render_stripes(s, cam, lights, n_lights, obj_pos, W, H, 0, 0, W, H, stripe, rank, n_ranks, d_local_framebuffer, h_framebuffer,
                                      d_sync_slot, seq, stream, flags);
}

// gathering rank: frame `seq` of the slot is complete when all ranks have pushed it (waits ON THE DEVICE, in stream order)
int rt_frame_complete_device(rt_scene *s, void *d_sync_slot, uint32_t seq, uint32_t n_ranks, uint32_t release, void *stream) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_sync_slot || !stream || n_ranks == 0 || n_ranks > 32) return fail(s, RT_ERR_INVALID, "rt_frame_complete_device: bad argument");
    CU(s, cudaSetDevice(s->device));
    CU(s, rt_launch_frame_complete((uint32_t *)d_sync_slot, seq, (int)n_ranks, release != 0, s->sync_timeout_ms, (cudaStream_t)stream));
    return RT_OK;
}

int rt_frame_release_device(rt_scene *s, void *d_sync_slot, uint32_t seq, void *stream) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_sync_slot || !stream) return fail(s, RT_ERR_INVALID, "rt_frame_release_device: bad argument");
    CU(s, cudaSetDevice(s->device));
    CU(s, rt_launch_frame_release((uint32_t *)d_sync_slot, seq, (cudaStream_t)stream));
    return RT_OK;
}

// ---- device buffers shareable between the per-GPU processes of one box (CUDA IPC over NVLink / NVSwitch) ----
int rt_device_alloc(int device, size_t bytes, void **out) {
    API_ENTER();
    if (!out || bytes == 0) return fail(nullptr, RT_ERR_INVALID, "rt_device_alloc: bad argument");
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaMalloc(out, bytes));
    return RT_OK;
}
int rt_device_free(int device, void *p) {
    API_ENTER();
    if (!p) return RT_OK;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaFree(p));
    return RT_OK;
}
int rt_ipc_export(void *d_ptr, unsigned char handle[64]) {
    API_ENTER();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle) return fail(nullptr, RT_ERR_INVALID, "rt_ipc_export: NULL argument");
    cudaIpcMemHandle_t h;
    CU(nullptr, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return RT_OK;
}
int rt_ipc_open(int device, const unsigned char handle[64], void **out) {
    API_ENTER();
    if (!handle || !out) return fail(nullptr, RT_ERR_INVALID, "rt_ipc_open: NULL argument");
    CU(nullptr, cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(nullptr, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
int rt_ipc_close(int device, void *p) {
    API_ENTER();
    if (!p) return RT_OK;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaIpcCloseMemHandle(p));
    return RT_OK;
}

int rt_untile_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t n_ranks, const void *d_all_tiles,
                     void *d_out_rgb, void *stream, uint32_t flags) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_all_tiles || !d_out_rgb || !tile_w || !tile_h || !n_ranks || !W || !H) return fail(s, RT_ERR_INVALID, "rt_untile_device: bad argument");
    if (tile_w % 8 || tile_h % 4 || (size_t)tile_w * tile_h * 3 > 48 * 1024) return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4 and at most 16384 pixels");
    Lane &L = *s->lanes[0];
    std::lock_guard<std::mutex> lk(L.mu);
    CU(s, cudaSetDevice(s->device));
    const uint32_t tiles_x = (W + tile_w - 1) / tile_w, tiles_y = (H + tile_h - 1) / tile_h;
    const uint32_t n_tiles = tiles_x * tiles_y;
    const uint32_t per_rank_max = (n_tiles + n_ranks - 1) / n_ranks;
    cudaStream_t st = stream ? (cudaStream_t)stream : L.stream;
    CU(s, rt_launch_untile((const uint8_t *)d_all_tiles, (uint8_t *)d_out_rgb, (int)W, (int)H, (int)tile_w, (int)tile_h, (int)n_ranks,
                           (int)per_rank_max, (flags & RT_LAYOUT_ROWMAJOR) != 0, st));
    if (!stream) CU(s, cudaStreamSynchronize(st));
    return RT_OK;
}

int rt_get_stats(rt_scene *s, rt_stats *out) {
    API_ENTER();
    if (!s || !out) return fail(s, RT_ERR_INVALID, "NULL argument");
    Lane &L = *s->lanes[s->last_lane.load()];
    std::lock_guard<std::mutex> lk(L.mu);
    CU(s, cudaSetDevice(s->device));
    if (L.stats_pending) { CU(s, cudaStreamSynchronize(L.last_stream)); collect_stats(s, L); }
    *out = L.stats;
    return RT_OK;
}

int rt_get_bvh_info(rt_scene *s, uint32_t mesh, rt_bvh_info *out) {
    API_ENTER();
    if (!s || !out) return fail(s, RT_ERR_INVALID, "NULL argument");
    if (mesh >= s->meshes.size()) return fail(s, RT_ERR_INVALID, "mesh index out of range");
    *out = s->meshes[mesh].info;
    return RT_OK;
}

int rt_synchronize(rt_scene *s) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    CU(s, cudaSetDevice(s->device));
    for (auto &lp : s->lanes) {
        std::lock_guard<std::mutex> lk(lp->mu);
        CU(s, cudaStreamSynchronize(lp->stream));
        if (lp->stats_pending) { CU(s, cudaStreamSynchronize(lp->last_stream)); collect_stats(s, *lp); }
    }
    return RT_OK;
}

int rt_cancel(rt_scene *s) {
    API_ENTER();
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    CU(s, cudaSetDevice(s->device));
    std::lock_guard<std::mutex> lk(s->mu);
    const unsigned int gen = s->cancel_gen.fetch_add(1u) + 1u;
    // tell the kernels in flight (they poll this word; calls that start from now on carry the new generation and ignore it)
    *s->h_cancel_word = gen;
    for (auto &lp : s->lanes)
        if (lp->d_counters)
            CU(s, cudaMemcpyAsync(lp->d_counters + RT_CNT_CANCEL, s->h_cancel_word, sizeof(unsigned int), cudaMemcpyHostToDevice, s->cancel_stream));
    CU(s, cudaStreamSynchronize(s->cancel_stream));
    return RT_OK;
}

int rt_measure_fma_peak(int device, int fp64, double *out_tflops) {
    API_ENTER();
    if (!out_tflops) return fail(nullptr, RT_ERR_INVALID, "NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(nullptr, RT_ERR_CUDA, "no such CUDA device"); }
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    CU(nullptr, rt_fma_peak(fp64 != 0, prop.multiProcessorCount, out_tflops));
    return RT_OK;
}

void *rt_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void rt_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// Pins caller-owned host memory (typically a POSIX shared-memory mapping that every per-GPU process of the box has mapped)
// and returns the address kernels of `device` use for it: each GPU then stores its tiles straight into the host
// framebuffer over ITS OWN PCIe link (rt_render_push_tiles with d_framebuffer / d_sync_slot = such addresses).
int rt_host_register(int device, void *ptr, size_t bytes, void **d_ptr) {
    API_ENTER();
    if (!ptr || !bytes || !d_ptr) return fail(nullptr, RT_ERR_INVALID, "rt_host_register: bad argument");
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    cudaError_t e = cudaHostGetDevicePointer(d_ptr, ptr, 0);
    if (e != cudaSuccess) { cudaHostUnregister(ptr); return fail(nullptr, RT_ERR_CUDA, cudaGetErrorString(e)); }
    return RT_OK;
}
int rt_host_unregister(int device, void *ptr) {
    API_ENTER();
    if (!ptr) return RT_OK;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaHostUnregister(ptr));
    return RT_OK;
}

}  // extern "C"
