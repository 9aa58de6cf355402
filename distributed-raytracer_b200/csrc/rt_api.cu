// rt_api.cu -- implementation of the C ABI declared in include/rt_b200.h.
//
// Lifecycle mirrors the reference worker (paths under /root/reference/):
//   rt_scene_create   <- register():        worker/distributed/main.go:101-129 (immutable meshes arrive once)
//   rt_frame_set      <- BulkTrace diff:    worker/distributed/main.go:56-64  (camera, lights, object positions)
//   rt_render_region  <- BulkTrace loop:    worker/distributed/main.go:67-88  /  draw(): worker/sequential/main.go:21-28
// No CPU fallback exists: every path below needs a CUDA device and the sm_100a kernels.
#include "rt_internal.h"
#include "go_math.h"
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>

struct rt_scene {
    std::mutex mu;
    std::string err;
    int device = 0;
    uint32_t precision = RT_FP32_FAST;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int max_leaf = 4;
    bool top_in_smem = true;
    int refill_below = 24;
    int shade_minb = 8;
    bool primary_persistent = false;

    std::vector<MeshDevice> meshes;
    std::vector<MeshView> h_views;
    MeshView *d_meshes = nullptr;
    std::vector<ObjRec> h_objs;
    ObjRec *d_objs = nullptr;
    LightRec *d_lights = nullptr;
    uint32_t n_lights = 0, cap_lights = 0;
    FrameParams frame;
    InlineFrame inline_frame;
    bool have_frame = false;

    // scratch
    HitRec *d_hits = nullptr; double *d_hit_colour = nullptr; uint32_t *d_hit_lit = nullptr; size_t cap_rays = 0;
    double *d_hit_geom = nullptr; bool split_shade = false;
    unsigned int *d_counters = nullptr; Counters *d_stats = nullptr;
    uint8_t *d_rgb = nullptr; size_t cap_rgb = 0;
    double *d_dbg_colour = nullptr; int32_t *d_dbg_face = nullptr, *d_dbg_obj = nullptr; uint32_t *d_dbg_lit = nullptr;
    size_t cap_dbg = 0;
    uint8_t *h_pinned = nullptr; size_t cap_pinned = 0;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaStream_t last_stream = nullptr;
    bool stats_pending = false;
    rt_stats stats;
    std::atomic<int> cancel{ 0 };
};

namespace {

thread_local std::string g_err;

int fail(rt_scene *s, int code, const std::string &msg) {
    if (s) s->err = msg;
    g_err = msg;
    return code;
}
#define CU(s, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail((s), e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T> cudaError_t upload(T **d, const T *h, size_t n, cudaStream_t st) {
    cudaError_t e = cudaMalloc((void **)d, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(*d, h, sizeof(T) * n, cudaMemcpyHostToDevice, st);
    return e;
}

void free_mesh(MeshDevice &m) {
    cudaFree(m.nodes); cudaFree(m.tri32); cudaFree(m.tri64); cudaFree(m.fverts); cudaFree(m.fnormals); cudaFree(m.fmat);
    cudaFree(m.materials);
    m = MeshDevice();
}

void destroy_locked(rt_scene *s) {
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto &m : s->meshes) free_mesh(m);
    cudaFree(s->d_meshes); cudaFree(s->d_objs); cudaFree(s->d_lights);
    cudaFree(s->d_hits); cudaFree(s->d_hit_colour); cudaFree(s->d_hit_lit); cudaFree(s->d_hit_geom); cudaFree(s->d_counters); cudaFree(s->d_stats);
    cudaFree(s->d_rgb); cudaFree(s->d_dbg_colour); cudaFree(s->d_dbg_face); cudaFree(s->d_dbg_obj); cudaFree(s->d_dbg_lit);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (auto &e : s->ev) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// gather one mesh into per-face arrays (host), validate indices
int flatten_mesh(const rt_mesh_desc &d, std::vector<double> &fverts, std::vector<double> &fnormals, std::vector<uint32_t> &fmat,
                 std::string &err) {
    const size_t nf = d.n_faces;
    if (nf && (!d.faces || !d.verts)) { err = "mesh: faces/verts pointer is NULL"; return RT_ERR_INVALID; }
    if (d.n_materials == 0 && nf) { err = "mesh: no materials"; return RT_ERR_INVALID; }
    if (d.n_materials && !d.materials) { err = "mesh: materials pointer is NULL"; return RT_ERR_INVALID; }
    if (d.n_normals && !d.normals) { err = "mesh: normals pointer is NULL"; return RT_ERR_INVALID; }
    fverts.resize(9 * nf);
    fmat.resize(nf);
    if (d.n_normals) fnormals.resize(9 * nf); else fnormals.clear();
    for (size_t f = 0; f < nf; f++) {
        const uint32_t *fc = d.faces + 7 * f;
        for (int v = 0; v < 3; v++) {
            if (fc[v] >= d.n_verts) { err = "mesh: vertex index out of range"; return RT_ERR_INVALID; }
            for (int k = 0; k < 3; k++) fverts[9 * f + 3 * v + k] = d.verts[3 * (size_t)fc[v] + k];
            if (d.n_normals) {
                if (fc[3 + v] >= d.n_normals) { err = "mesh: normal index out of range"; return RT_ERR_INVALID; }
                for (int k = 0; k < 3; k++) fnormals[9 * f + 3 * v + k] = d.normals[3 * (size_t)fc[3 + v] + k];
            }
        }
        if (fc[6] >= d.n_materials) { err = "mesh: material index out of range"; return RT_ERR_INVALID; }
        fmat[f] = fc[6];
    }
    return RT_OK;
}

int ensure_scratch(rt_scene *s, size_t nrays, size_t out_bytes, bool need_rgb, bool dbg, size_t npix_out) {
    if (nrays > s->cap_rays) {
        cudaFree(s->d_hits); cudaFree(s->d_hit_colour); cudaFree(s->d_hit_lit); cudaFree(s->d_hit_geom);
        s->d_hits = nullptr; s->d_hit_colour = nullptr; s->d_hit_lit = nullptr; s->d_hit_geom = nullptr; s->cap_rays = 0;
        CU(s, cudaMalloc((void **)&s->d_hits, sizeof(HitRec) * nrays));
        CU(s, cudaMalloc((void **)&s->d_hit_colour, sizeof(double) * 3 * nrays));
        CU(s, cudaMalloc((void **)&s->d_hit_lit, sizeof(uint32_t) * nrays));
        if (s->split_shade) CU(s, cudaMalloc((void **)&s->d_hit_geom, sizeof(double) * 6 * nrays));
        s->cap_rays = nrays;
    }
    if (need_rgb && out_bytes > s->cap_rgb) {
        cudaFree(s->d_rgb); s->d_rgb = nullptr; s->cap_rgb = 0;
        CU(s, cudaMalloc((void **)&s->d_rgb, out_bytes));
        s->cap_rgb = out_bytes;
    }
    if (dbg && npix_out > s->cap_dbg) {
        cudaFree(s->d_dbg_colour); cudaFree(s->d_dbg_face); cudaFree(s->d_dbg_obj); cudaFree(s->d_dbg_lit);
        s->d_dbg_colour = nullptr; s->d_dbg_face = nullptr; s->d_dbg_obj = nullptr; s->d_dbg_lit = nullptr; s->cap_dbg = 0;
        CU(s, cudaMalloc((void **)&s->d_dbg_colour, sizeof(double) * 3 * npix_out));
        CU(s, cudaMalloc((void **)&s->d_dbg_face, sizeof(int32_t) * npix_out));
        CU(s, cudaMalloc((void **)&s->d_dbg_obj, sizeof(int32_t) * npix_out));
        CU(s, cudaMalloc((void **)&s->d_dbg_lit, sizeof(uint32_t) * npix_out));
        s->cap_dbg = npix_out;
    }
    return RT_OK;
}

int frame_set_locked(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos) {
    if (!cam) return fail(s, RT_ERR_INVALID, "rt_frame_set: camera is NULL");
    if (n_lights && !lights) return fail(s, RT_ERR_INVALID, "rt_frame_set: lights is NULL");
    CU(s, cudaSetDevice(s->device));
    FrameParams &fp = s->frame;
    for (int k = 0; k < 3; k++) { fp.cam_pos[k] = cam->pos[k]; fp.fwd[k] = cam->forward[k]; fp.left[k] = cam->left[k]; fp.up[k] = cam->up[k]; }
    fp.tan_half = std::isnan(cam->tan_half_fov) ? rt_go_tan(cam->fov / 2.0) : cam->tan_half_fov;   // tracer.go:17
    fp.n_lights = (int)n_lights;
    fp.n_obj = (int)s->h_objs.size();
    s->n_lights = n_lights;
    if (obj_pos)
        for (size_t k = 0; k < s->h_objs.size(); k++)
            for (int c = 0; c < 3; c++) s->h_objs[k].pos[c] = obj_pos[3 * k + c];
    InlineFrame &fr = s->inline_frame;
    fr.use_lights = n_lights <= RT_INLINE_LIGHTS;
    fr.use_objs = s->h_objs.size() <= RT_INLINE_OBJS;
    static_assert(sizeof(LightRec) == sizeof(rt_light), "light layout");
    if (fr.use_lights && n_lights) memcpy(fr.lights, lights, sizeof(LightRec) * n_lights);
    if (fr.use_objs && !s->h_objs.empty()) memcpy(fr.objs, s->h_objs.data(), sizeof(ObjRec) * s->h_objs.size());
    if (!fr.use_lights || (!fr.use_objs && obj_pos)) {
        // large light / object sets live in device memory: the copies must not overtake a render still in flight
        CU(s, cudaStreamSynchronize(s->stream));
        if (s->last_stream) CU(s, cudaStreamSynchronize(s->last_stream));
        if (!fr.use_lights) {
            if (n_lights > s->cap_lights) {
                cudaFree(s->d_lights); s->d_lights = nullptr; s->cap_lights = 0;
                CU(s, cudaMalloc((void **)&s->d_lights, sizeof(LightRec) * n_lights));
                s->cap_lights = n_lights;
            }
            CU(s, cudaMemcpyAsync(s->d_lights, lights, sizeof(LightRec) * n_lights, cudaMemcpyHostToDevice, s->stream));
        }
        if (!fr.use_objs && obj_pos)
            CU(s, cudaMemcpyAsync(s->d_objs, s->h_objs.data(), sizeof(ObjRec) * s->h_objs.size(), cudaMemcpyHostToDevice, s->stream));
        CU(s, cudaStreamSynchronize(s->stream));   // `lights` is caller memory: do not retain it
    }
    s->have_frame = true;
    return RT_OK;
}

struct RenderOut {
    uint8_t *h_rgb = nullptr; double *h_colour = nullptr; int32_t *h_face = nullptr, *h_obj = nullptr; uint32_t *h_lit = nullptr;
    uint8_t *d_rgb = nullptr;      // external device output (device variants)
    cudaStream_t ext_stream = nullptr; bool async = false;
};

void collect_stats(rt_scene *s) {
    if (!s->stats_pending) return;
    unsigned int counters[4] = { 0, 0, 0, 0 };
    cudaMemcpy(counters, s->d_counters, sizeof counters, cudaMemcpyDeviceToHost);
    Counters c[2];
    memset(c, 0, sizeof c);
    cudaMemcpy(c, s->d_stats, sizeof c, cudaMemcpyDeviceToHost);
    s->stats.hit_pixels = counters[1];
    s->stats.shadow_rays = (uint64_t)counters[1] * (uint64_t)s->frame.n_lights;
    s->stats.nodes_primary = c[0].nodes; s->stats.tris_primary = c[0].tris;
    s->stats.nodes_shade = c[1].nodes; s->stats.tris_shade = c[1].tris;
    s->stats.nodes_fetched = c[0].nodes + c[1].nodes;
    s->stats.tris_fetched = c[0].tris + c[1].tris;
    s->stats.node_bytes = RT_NODE_BYTES;
    s->stats.tri_bytes = s->precision == RT_FP64_EXACT ? RT_TRI64_BYTES : RT_TRI32_BYTES;
    cudaEventElapsedTime(&s->stats.ms_primary, s->ev[0], s->ev[1]);
    cudaEventElapsedTime(&s->stats.ms_shade, s->ev[1], s->ev[2]);
    cudaEventElapsedTime(&s->stats.ms_write, s->ev[2], s->ev[3]);
    cudaEventElapsedTime(&s->stats.ms_total, s->ev[0], s->ev[3]);
    s->stats_pending = false;
}

// the whole render path: clear -> K2 -> K4 -> K5 -> (copy out)
int render_locked(rt_scene *s, const FrameParams &fp, uint32_t flags, size_t npix_out, const RenderOut &out) {
    if (!s->have_frame) return fail(s, RT_ERR_STATE, "render before rt_frame_set");
    CU(s, cudaSetDevice(s->device));
    s->cancel.store(0);   // stats of earlier frames are collected lazily (rt_get_stats): no per-frame host sync
    const size_t nrays = ray_count(fp);
    const size_t out_bytes = npix_out * 3;
    const bool dbg = out.h_colour || out.h_face || out.h_obj || out.h_lit;
    const bool internal_rgb = out.d_rgb == nullptr;
    int rc = ensure_scratch(s, nrays, out_bytes, internal_rgb, dbg, npix_out);
    if (rc != RT_OK) return rc;
    cudaStream_t st = out.ext_stream ? out.ext_stream : s->stream;
    uint8_t *d_rgb = internal_rgb ? s->d_rgb : out.d_rgb;

    RenderJob job;
    memset(&job, 0, sizeof job);
    job.fp = fp;
    job.d_meshes = s->d_meshes; job.d_objs = s->d_objs; job.d_lights = s->d_lights;
    job.fr = s->inline_frame;
    job.precision = s->precision; job.flags = flags;
    job.d_hits = s->d_hits; job.d_hit_colour = s->d_hit_colour; job.d_hit_lit = s->d_hit_lit;
    job.d_hit_geom = s->d_hit_geom; job.split_shade = s->split_shade;
    job.d_counters = s->d_counters; job.d_stats = s->d_stats;
    job.d_out_rgb = d_rgb;
    job.d_out_colour = out.h_colour ? s->d_dbg_colour : nullptr;
    job.d_out_face = out.h_face ? s->d_dbg_face : nullptr;
    job.d_out_obj = out.h_obj ? s->d_dbg_obj : nullptr;
    job.d_out_lit = out.h_lit ? s->d_dbg_lit : nullptr;
    job.sm_count = s->sm_count;
    job.single_mesh = s->h_objs.size() == 1;
    if (job.single_mesh) {
        job.h_mesh0 = s->h_views[s->h_objs[0].mesh];
        for (int k = 0; k < 3; k++) job.h_obj0_pos[k] = s->h_objs[0].pos[k];
    }
    job.top_in_smem = s->top_in_smem;
    job.refill_below = s->refill_below;
    job.shade_minb = s->shade_minb;
    job.primary_persistent = s->primary_persistent;

    CU(s, cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned int), st));
    CU(s, cudaMemsetAsync(s->d_stats, 0, 2 * sizeof(Counters), st));
    // background: sequential/main.go:17 surface.FillRect(nil, 0); distributed/main.go:70 default 0,0,0
    if (fp.tile_final) CU(s, rt_launch_clear_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, d_rgb, st));   // only this rank's pixels (peer framebuffer)
    else CU(s, cudaMemsetAsync(d_rgb, 0, out_bytes, st));
    if (job.d_out_colour) CU(s, cudaMemsetAsync(job.d_out_colour, 0, sizeof(double) * 3 * npix_out, st));
    if (job.d_out_face) CU(s, cudaMemsetAsync(job.d_out_face, 0xff, sizeof(int32_t) * npix_out, st));
    if (job.d_out_obj) CU(s, cudaMemsetAsync(job.d_out_obj, 0xff, sizeof(int32_t) * npix_out, st));
    if (job.d_out_lit) CU(s, cudaMemsetAsync(job.d_out_lit, 0, sizeof(uint32_t) * npix_out, st));

    CU(s, cudaEventRecord(s->ev[0], st));
    CU(s, rt_launch_primary(job, st));
    CU(s, cudaEventRecord(s->ev[1], st));
    if (s->cancel.load()) { cudaStreamSynchronize(st); return fail(s, RT_ERR_CANCELLED, "cancelled"); }
    CU(s, rt_launch_shade(job, st));
    CU(s, cudaEventRecord(s->ev[2], st));
    CU(s, rt_launch_write(job, st));
    CU(s, cudaEventRecord(s->ev[3], st));

    s->stats.primary_rays = (uint64_t)fp.w * (uint64_t)fp.h;
    if (fp.tile_w > 0) s->stats.primary_rays = nrays;
    s->stats_pending = true;
    s->last_stream = st;
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
            if (out_bytes > s->cap_pinned) {
                if (s->h_pinned) cudaFreeHost(s->h_pinned);
                s->h_pinned = nullptr; s->cap_pinned = 0;
                CU(s, cudaMallocHost((void **)&s->h_pinned, out_bytes));
                s->cap_pinned = out_bytes;
            }
            CU(s, cudaMemcpyAsync(s->h_pinned, d_rgb, out_bytes, cudaMemcpyDeviceToHost, st));
            CU(s, cudaStreamSynchronize(st));
            memcpy(out.h_rgb, s->h_pinned, out_bytes);
        }
    }
    if (out.h_colour) CU(s, cudaMemcpyAsync(out.h_colour, job.d_out_colour, sizeof(double) * 3 * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_face) CU(s, cudaMemcpyAsync(out.h_face, job.d_out_face, sizeof(int32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_obj) CU(s, cudaMemcpyAsync(out.h_obj, job.d_out_obj, sizeof(int32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    if (out.h_lit) CU(s, cudaMemcpyAsync(out.h_lit, job.d_out_lit, sizeof(uint32_t) * npix_out, cudaMemcpyDeviceToHost, st));
    CU(s, cudaStreamSynchronize(st));
    // counters and event times are read lazily by rt_get_stats: two blocking device->host copies per call would be on
    // the critical path of concurrent BulkTrace callers
    return RT_OK;
}

int region_params(rt_scene *s, FrameParams &fp, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
    if (W == 0 || H == 0 || w == 0 || h == 0) return fail(s, RT_ERR_INVALID, "empty screen or region");
    if (W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "screen larger than 65536 pixels on a side");
    if ((uint64_t)x + w > W || (uint64_t)y + h > H) return fail(s, RT_ERR_INVALID, "region exceeds the screen");
    fp = s->frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = (int)x; fp.y = (int)y; fp.w = (int)w; fp.h = (int)h;
    fp.tile_w = fp.tile_h = fp.tiles_x = fp.rank = fp.n_ranks = fp.n_local_tiles = fp.tile_final = 0;
    return RT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *rt_last_error(const rt_scene *scene) { return scene ? scene->err.c_str() : g_err.c_str(); }

int rt_scene_create(const rt_scene_desc *desc, rt_scene **out) {
    if (!desc || !out) return fail(nullptr, RT_ERR_INVALID, "rt_scene_create: NULL argument");
    *out = nullptr;
    if (desc->precision != RT_FP64_EXACT && desc->precision != RT_FP32_FAST) return fail(nullptr, RT_ERR_INVALID, "unknown precision mode");
    if (desc->n_meshes && !desc->meshes) return fail(nullptr, RT_ERR_INVALID, "meshes is NULL");
    if (desc->n_objects && (!desc->obj_pos || !desc->obj_mesh)) return fail(nullptr, RT_ERR_INVALID, "objects are NULL");
    for (uint32_t k = 0; k < desc->n_objects; k++)
        if (desc->obj_mesh[k] >= desc->n_meshes) return fail(nullptr, RT_ERR_INVALID, "object refers to a mesh out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, RT_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
    }
    if (desc->device < 0 || desc->device >= ndev) return fail(nullptr, RT_ERR_INVALID, "device ordinal out of range");
    rt_scene *s = new (std::nothrow) rt_scene();
    if (!s) return fail(nullptr, RT_ERR_OOM, "out of host memory");
    std::unique_lock<std::mutex> lk(s->mu);
    s->device = desc->device;
    s->precision = desc->precision;
    s->max_leaf = env_int("RT_B200_MAX_LEAF", 4);
    s->shade_minb = env_int("RT_B200_SHADE_MINB", 8);
    s->primary_persistent = env_int("RT_B200_PERSISTENT", 0) != 0;
    s->split_shade = env_int("RT_B200_SPLIT_SHADE", 0) != 0;
    s->top_in_smem = env_int("RT_B200_TOP_SMEM", 0) != 0;
    s->refill_below = env_int("RT_B200_REFILL", 1);
    memset(&s->stats, 0, sizeof s->stats);
    memset(&s->frame, 0, sizeof s->frame);
    memset(&s->inline_frame, 0, sizeof s->inline_frame);
    int rc = RT_OK;
    auto bail = [&](int code) { lk.unlock(); std::string e = s->err; destroy_locked(s); delete s; g_err = e; return code; };
#define CUB_(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(s, e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return bail(e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA); } } while (0)
    CUB_(cudaSetDevice(s->device));
    {
        cudaDeviceProp prop;
        CUB_(cudaGetDeviceProperties(&prop, s->device));
        s->sm_count = prop.multiProcessorCount;
        // the kernels are compiled for sm_100a only: on any other device the first launch fails with
        // cudaErrorNoKernelImageForDevice and the error is returned (no fallback)
    }
    CUB_(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto &e : s->ev) CUB_(cudaEventCreate(&e));
    CUB_(cudaMalloc((void **)&s->d_counters, 4 * sizeof(unsigned int)));
    CUB_(cudaMalloc((void **)&s->d_stats, 2 * sizeof(Counters)));

    s->meshes.resize(desc->n_meshes);
    s->h_views.resize(desc->n_meshes);
    for (uint32_t k = 0; k < desc->n_meshes; k++) {
        const rt_mesh_desc &d = desc->meshes[k];
        MeshDevice &m = s->meshes[k];
        std::vector<double> fverts, fnormals; std::vector<uint32_t> fmat;
        std::string e;
        rc = flatten_mesh(d, fverts, fnormals, fmat, e);
        if (rc != RT_OK) { fail(s, rc, e); return bail(rc); }
        m.n_faces = d.n_faces; m.n_materials = d.n_materials;
        CUB_(upload(&m.fverts, fverts.data(), fverts.size(), s->stream));
        if (!fnormals.empty()) CUB_(upload(&m.fnormals, fnormals.data(), fnormals.size(), s->stream));
        CUB_(upload(&m.fmat, fmat.data(), fmat.size(), s->stream));
        CUB_(upload(&m.materials, d.materials, (size_t)10 * d.n_materials, s->stream));
        CUB_(cudaStreamSynchronize(s->stream));
        if (!rt_build_bvh(m, s->precision == RT_FP32_FAST, s->precision == RT_FP64_EXACT, s->max_leaf, s->stream, e)) {
            fail(s, RT_ERR_CUDA, "BVH build: " + e);
            return bail(RT_ERR_CUDA);
        }
        MeshView &v = s->h_views[k];
        v.nodes = m.nodes; v.tri32 = m.tri32; v.tri64 = m.tri64; v.fverts = m.fverts; v.fnormals = m.fnormals; v.fmat = m.fmat;
        v.materials = m.materials; v.n_faces = m.n_faces; v.n_nodes = m.n_nodes; v.root_ref = m.root_ref;
        for (int c = 0; c < 3; c++) { v.lo[c] = m.lo[c]; v.hi[c] = m.hi[c]; }
    }
    CUB_(upload(&s->d_meshes, s->h_views.data(), s->h_views.size(), s->stream));
    s->h_objs.resize(desc->n_objects);
    for (uint32_t k = 0; k < desc->n_objects; k++) {
        for (int c = 0; c < 3; c++) s->h_objs[k].pos[c] = desc->obj_pos[3 * (size_t)k + c];
        s->h_objs[k].mesh = desc->obj_mesh[k]; s->h_objs[k].pad = 0;
    }
    CUB_(upload(&s->d_objs, s->h_objs.data(), s->h_objs.size(), s->stream));
    CUB_(cudaStreamSynchronize(s->stream));
#undef CUB_
    lk.unlock();
    *out = s;
    return RT_OK;
}

void rt_scene_destroy(rt_scene *scene) {
    if (!scene) return;
    { std::lock_guard<std::mutex> lk(scene->mu); destroy_locked(scene); }
    delete scene;
}

int rt_frame_set(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    std::lock_guard<std::mutex> lk(s->mu);
    return frame_set_locked(s, cam, lights, n_lights, obj_pos);
}

int rt_render_region_debug(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                           double *out_colour, int32_t *out_face, int32_t *out_obj, uint32_t *out_lit, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    std::lock_guard<std::mutex> lk(s->mu);
    FrameParams fp;
    int rc = region_params(s, fp, W, H, x, y, w, h);
    if (rc != RT_OK) return rc;
    RenderOut o; o.h_rgb = out_rgb; o.h_colour = out_colour; o.h_face = out_face; o.h_obj = out_obj; o.h_lit = out_lit;
    return render_locked(s, fp, flags, (size_t)w * h, o);
}

int rt_render_region(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb,
                     uint32_t flags) {
    if (!out_rgb) return fail(s, RT_ERR_INVALID, "out_rgb is NULL");
    return rt_render_region_debug(s, W, H, x, y, w, h, out_rgb, nullptr, nullptr, nullptr, nullptr, flags);
}

int rt_bulk_trace(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                  uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!out_rgb) return fail(s, RT_ERR_INVALID, "out_rgb is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = frame_set_locked(s, cam, lights, n_lights, obj_pos);
    if (rc != RT_OK) return rc;
    FrameParams fp;
    rc = region_params(s, fp, W, H, x, y, w, h);
    if (rc != RT_OK) return rc;
    RenderOut o; o.h_rgb = out_rgb;
    return render_locked(s, fp, flags, (size_t)w * h, o);
}

int rt_render_region_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void *d_out_rgb,
                            void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_out_rgb) return fail(s, RT_ERR_INVALID, "d_out_rgb is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    FrameParams fp;
    int rc = region_params(s, fp, W, H, x, y, w, h);
    if (rc != RT_OK) return rc;
    RenderOut o; o.d_rgb = (uint8_t *)d_out_rgb; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    rc = render_locked(s, fp, flags, (size_t)w * h, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(s->stream)); collect_stats(s); }
    return rc;
}

int rt_render_tiles_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                           void *d_out_tiles, void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_out_tiles) return fail(s, RT_ERR_INVALID, "d_out_tiles is NULL");
    if (W == 0 || H == 0 || W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "bad screen size");
    if (tile_w == 0 || tile_h == 0 || tile_w % 8 || tile_h % 4) return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4");
    if (n_ranks == 0 || rank >= n_ranks) return fail(s, RT_ERR_INVALID, "bad rank / n_ranks");
    std::lock_guard<std::mutex> lk(s->mu);
    FrameParams fp = s->frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = 0; fp.y = 0; fp.w = (int)W; fp.h = (int)H;
    fp.tile_w = (int)tile_w; fp.tile_h = (int)tile_h;
    fp.tiles_x = (int)((W + tile_w - 1) / tile_w);
    const uint32_t tiles_y = (H + tile_h - 1) / tile_h;
    const uint32_t n_tiles = (uint32_t)fp.tiles_x * tiles_y;
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)((n_tiles > rank) ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    fp.tile_final = 0;
    RenderOut o; o.d_rgb = (uint8_t *)d_out_tiles; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    int rc = render_locked(s, fp, flags & ~RT_LAYOUT_ROWMAJOR, (size_t)fp.n_local_tiles * tile_w * tile_h, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(s->stream)); collect_stats(s); }
    return rc;
}

int rt_render_tiles_peer(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                         void *d_framebuffer, void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_framebuffer) return fail(s, RT_ERR_INVALID, "d_framebuffer is NULL");
    if (W == 0 || H == 0 || W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "bad screen size");
    if (tile_w == 0 || tile_h == 0 || tile_w % 8 || tile_h % 4) return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4");
    if (n_ranks == 0 || rank >= n_ranks) return fail(s, RT_ERR_INVALID, "bad rank / n_ranks");
    std::lock_guard<std::mutex> lk(s->mu);
    FrameParams fp = s->frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = 0; fp.y = 0; fp.w = (int)W; fp.h = (int)H;
    fp.tile_w = (int)tile_w; fp.tile_h = (int)tile_h;
    fp.tiles_x = (int)((W + tile_w - 1) / tile_w);
    const uint32_t tiles_y = (H + tile_h - 1) / tile_h;
    const uint32_t n_tiles = (uint32_t)fp.tiles_x * tiles_y;
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)((n_tiles > rank) ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    fp.tile_final = 1;
    RenderOut o; o.d_rgb = (uint8_t *)d_framebuffer; o.ext_stream = (cudaStream_t)stream; o.async = stream != nullptr;
    int rc = render_locked(s, fp, flags, (size_t)W * H, o);
    if (rc == RT_OK && !o.async) { CU(s, cudaStreamSynchronize(s->stream)); collect_stats(s); }
    return rc;
}

int rt_push_tiles_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks,
                         const void *d_tiles, void *d_framebuffer, void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_tiles || !d_framebuffer || !W || !H) return fail(s, RT_ERR_INVALID, "rt_push_tiles_device: bad argument");
    if (tile_w == 0 || tile_h == 0 || tile_w % 8 || tile_h % 4 || (size_t)tile_w * tile_h * 3 > 48 * 1024)
        return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4 and at most 16384 pixels");
    if (n_ranks == 0 || rank >= n_ranks) return fail(s, RT_ERR_INVALID, "bad rank / n_ranks");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(s, cudaSetDevice(s->device));
    FrameParams fp;
    memset(&fp, 0, sizeof fp);
    fp.W = (int)W; fp.H = (int)H; fp.tile_w = (int)tile_w; fp.tile_h = (int)tile_h;
    fp.tiles_x = (int)((W + tile_w - 1) / tile_w);
    const uint32_t n_tiles = (uint32_t)fp.tiles_x * ((H + tile_h - 1) / tile_h);
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)((n_tiles > rank) ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    CU(s, rt_launch_push_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, (const uint8_t *)d_tiles, (uint8_t *)d_framebuffer, st));
    if (!stream) CU(s, cudaStreamSynchronize(st));
    return RT_OK;
}

// rt_frame_set + rt_render_tiles_device + rt_push_tiles_device in one call (one lock, one ctypes/cgo crossing per frame)
int rt_render_push_tiles(rt_scene *s, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                         uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t n_ranks, void *d_tiles, void *d_framebuffer,
                         void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_tiles || !d_framebuffer || !stream) return fail(s, RT_ERR_INVALID, "rt_render_push_tiles: NULL buffer or stream");
    if (W == 0 || H == 0 || W > 65536 || H > 65536) return fail(s, RT_ERR_INVALID, "bad screen size");
    if (tile_w == 0 || tile_h == 0 || tile_w % 8 || tile_h % 4 || (size_t)tile_w * tile_h * 3 > 48 * 1024)
        return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4 and at most 16384 pixels");
    if (n_ranks == 0 || rank >= n_ranks) return fail(s, RT_ERR_INVALID, "bad rank / n_ranks");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = frame_set_locked(s, cam, lights, n_lights, obj_pos);
    if (rc != RT_OK) return rc;
    FrameParams fp = s->frame;
    fp.W = (int)W; fp.H = (int)H; fp.x = 0; fp.y = 0; fp.w = (int)W; fp.h = (int)H;
    fp.tile_w = (int)tile_w; fp.tile_h = (int)tile_h;
    fp.tiles_x = (int)((W + tile_w - 1) / tile_w);
    const uint32_t n_tiles = (uint32_t)fp.tiles_x * ((H + tile_h - 1) / tile_h);
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)((n_tiles > rank) ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    fp.tile_final = 0;
    RenderOut o; o.d_rgb = (uint8_t *)d_tiles; o.ext_stream = (cudaStream_t)stream; o.async = true;
    rc = render_locked(s, fp, flags & ~RT_LAYOUT_ROWMAJOR, (size_t)fp.n_local_tiles * tile_w * tile_h, o);
    if (rc != RT_OK) return rc;
    CU(s, rt_launch_push_tiles(fp, (flags & RT_LAYOUT_ROWMAJOR) != 0, (const uint8_t *)d_tiles, (uint8_t *)d_framebuffer, (cudaStream_t)stream));
    return RT_OK;
}

// ---- device buffers shareable between the per-GPU processes of one box (CUDA IPC over NVLink / NVSwitch) ----
int rt_device_alloc(int device, size_t bytes, void **out) {
    if (!out || bytes == 0) return fail(nullptr, RT_ERR_INVALID, "rt_device_alloc: bad argument");
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaMalloc(out, bytes));
    return RT_OK;
}
int rt_device_free(int device, void *p) {
    if (!p) return RT_OK;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaFree(p));
    return RT_OK;
}
int rt_ipc_export(void *d_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle) return fail(nullptr, RT_ERR_INVALID, "rt_ipc_export: NULL argument");
    cudaIpcMemHandle_t h;
    CU(nullptr, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return RT_OK;
}
int rt_ipc_open(int device, const unsigned char handle[64], void **out) {
    if (!handle || !out) return fail(nullptr, RT_ERR_INVALID, "rt_ipc_open: NULL argument");
    CU(nullptr, cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(nullptr, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
int rt_ipc_close(int device, void *p) {
    if (!p) return RT_OK;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaIpcCloseMemHandle(p));
    return RT_OK;
}

int rt_untile_device(rt_scene *s, uint32_t W, uint32_t H, uint32_t tile_w, uint32_t tile_h, uint32_t n_ranks, const void *d_all_tiles,
                     void *d_out_rgb, void *stream, uint32_t flags) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    if (!d_all_tiles || !d_out_rgb || !tile_w || !tile_h || !n_ranks || !W || !H) return fail(s, RT_ERR_INVALID, "rt_untile_device: bad argument");
    if (tile_w % 8 || tile_h % 4 || (size_t)tile_w * tile_h * 3 > 48 * 1024) return fail(s, RT_ERR_INVALID, "tile size must be a multiple of 8x4 and at most 16384 pixels");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(s, cudaSetDevice(s->device));
    const uint32_t tiles_x = (W + tile_w - 1) / tile_w, tiles_y = (H + tile_h - 1) / tile_h;
    const uint32_t n_tiles = tiles_x * tiles_y;
    const uint32_t per_rank_max = (n_tiles + n_ranks - 1) / n_ranks;
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    CU(s, rt_launch_untile((const uint8_t *)d_all_tiles, (uint8_t *)d_out_rgb, (int)W, (int)H, (int)tile_w, (int)tile_h, (int)n_ranks,
                           (int)per_rank_max, (flags & RT_LAYOUT_ROWMAJOR) != 0, st));
    if (!stream) CU(s, cudaStreamSynchronize(st));
    return RT_OK;
}

int rt_get_stats(rt_scene *s, rt_stats *out) {
    if (!s || !out) return fail(s, RT_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(s, cudaSetDevice(s->device));
    if (s->stats_pending) { CU(s, cudaStreamSynchronize(s->last_stream)); collect_stats(s); }
    *out = s->stats;
    return RT_OK;
}

int rt_get_bvh_info(rt_scene *s, uint32_t mesh, rt_bvh_info *out) {
    if (!s || !out) return fail(s, RT_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (mesh >= s->meshes.size()) return fail(s, RT_ERR_INVALID, "mesh index out of range");
    *out = s->meshes[mesh].info;
    return RT_OK;
}

int rt_synchronize(rt_scene *s) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    if (s->stats_pending) { CU(s, cudaStreamSynchronize(s->last_stream)); collect_stats(s); }
    return RT_OK;
}

int rt_cancel(rt_scene *s) {
    if (!s) return fail(nullptr, RT_ERR_INVALID, "NULL scene");
    s->cancel.store(1);
    return RT_OK;
}

int rt_measure_fma_peak(int device, int fp64, double *out_tflops) {
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

}  // extern "C"
