// render.cu -- the render kernels (sm_100a):
//   K2  k_primary_*        primary-ray generation + nearest-hit wide-BVH traversal, appends compact HitRecs
//   K3  Moller-Trumbore    device functions tri_test32 / tri_intersection_exact inlined in K2 and K4
//   K4  k_shade            fused shadow rays + Phong shading per primary hit (fp64, reference order)
//   K5  k_write            8-bit framebuffer writer (+ optional dense debug outputs)
//       k_untile           multi-GPU: interleaved-tile buffers -> framebuffer
//
// They replace the per-pixel double loop of worker/distributed/main.go:67-88 and
// worker/sequential/main.go:21-28 (each iteration = tracer.Trace, worker/shared/tracer/tracer.go:81-91).
#include "rt_internal.h"
#include "rt_traverse.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void flush_counters(Counters *g, const Counters &c) {
    if (g) {
        unsigned long long n = c.nodes, t = c.tris;
        // warp-aggregate
        for (int o = 16; o > 0; o >>= 1) { n += __shfl_down_sync(FULL, n, o); t += __shfl_down_sync(FULL, t, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&g->nodes, n); atomicAdd(&g->tris, t); }
    }
}

// lights / object positions come by value (a __grid_constant__ parameter may have its address taken) when they fit
__device__ __forceinline__ SceneView bind_frame(SceneView sc, const InlineFrame &fr, uint32_t *trav_smem = nullptr) {
    if (fr.use_lights) sc.lights = fr.lights;
    if (fr.use_objs) sc.objs = fr.objs;
    sc.trav_smem = trav_smem;
    sc.defer = sc.tune ? *(const volatile unsigned int *)(sc.tune + 2) : 0u;      // once per CTA
    return sc;
}
// the shared-memory part of the traversal stacks of a 128-thread CTA (SceneView::trav_smem): 8 KB
#if RT_SMEM_STACK
#define RT_DECLARE_TRAV_SMEM() __shared__ uint32_t s_trav[RT_SMEM_STACK_WORDS * RT_SMEM_STACK_STRIDE]; uint32_t *trav_smem = s_trav + threadIdx.x
#else
#define RT_DECLARE_TRAV_SMEM() uint32_t *trav_smem = nullptr
#endif

// rt_cancel (ctx.Err() of the reference loop, distributed/main.go:73): the generation in the lane's counter block has moved
// past the one this frame was launched with.  Wrap-safe; one L2-resident word.
__device__ __forceinline__ bool frame_cancelled(const unsigned int *counters, const FrameParams &fp) {
    return (int)(*(const volatile unsigned int *)(counters + RT_CNT_CANCEL) - fp.cancel_gen) > 0;
}
// the same, decided by lane 0 for the whole warp
__device__ __forceinline__ bool warp_cancelled(const unsigned int *counters, const FrameParams &fp) {
    int c = 0;
    if ((threadIdx.x & 31) == 0) c = frame_cancelled(counters, fp);
    return __shfl_sync(FULL, c, 0) != 0;
}

// append this lane's hit (if any) to the compact hit list; all 32 lanes must call
__device__ __forceinline__ void append_hit(bool hit, uint32_t ray, int obj, int face, float t, HitRec *hits,
                                           unsigned int *hit_count) {
    unsigned m = __ballot_sync(FULL, hit);
    if (m == 0) return;
    unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(hit_count, (unsigned)__popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (hit) {
        HitRec h; h.pix = ray; h.obj = obj; h.face = face; h.t = t;
        hits[base + __popc(m & ((1u << lane) - 1u))] = h;
    }
}

// Self-tuning of the any-hit postponement bound (SceneView::tune): every warp that shaded hits adds its shadow rays and how many of
// them were occluded; whenever the (sampled) count crosses a multiple of 16 384 the warp that crossed it re-derives the bound from the running
// ratio.  One CTA in four takes part (a sample is all a tuning knob needs).  All 32 lanes call; lanes without a hit are ignored.
__device__ __forceinline__ void tune_any_defer(unsigned int *tune, unsigned int *tune_host, bool valid, uint32_t lit, int n_lights) {
    if (!tune || (blockIdx.x & 3u)) return;
    unsigned tot = valid ? (unsigned)n_lights : 0u, occ = valid ? (unsigned)(n_lights - __popc(lit)) : 0u;
    tot = __reduce_add_sync(FULL, tot);
    occ = __reduce_add_sync(FULL, occ);
    if ((threadIdx.x & 31) == 0 && tot) {
        const unsigned old = atomicAdd(&tune[0], tot), oc = atomicAdd(&tune[1], occ) + occ, now = old + tot;
        if ((now >> 14) != (old >> 14)) {
            const float f = (float)oc / (float)now;
            const unsigned int bound = f < 0.15f ? 16u : (f < 0.40f ? 8u : 1u);
            *(volatile unsigned int *)(tune + 2) = bound;
            if (tune_host) *(volatile unsigned int *)tune_host = bound;
            if (now > (1u << 30)) { *(volatile unsigned int *)tune = 0u; *(volatile unsigned int *)(tune + 1) = 0u; }   // a fresh window; racing adds only blur a ratio
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2, FP64_EXACT: one thread per ray, 8x4 pixel block per warp
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128) k_primary_exact(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr, bool brute,
                                                        HitRec *hits, unsigned int *counters, Counters *gstats) {
    const SceneView sc = bind_frame(sc_in, fr);
    uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nrays = ray_count(fp);
    Counters c; c.nodes = 0; c.tris = 0;
    bool hit = false;
    TraceExact tr; tr.face = -1; tr.obj = -1;
    if (ray < nrays && !frame_cancelled(counters, fp)) {
        PixMap pm = map_ray(ray, fp, false);
        if (pm.valid && !(fp.tile_w > 0 && !brute && block_culled(fp, pm.i - (int)(ray & 7u), pm.j - (int)((ray >> 3) & 3u)))) {
            dv d = primary_dir_exact(pm.i, pm.j, fp);
            hit = trace_exact(sc, fp, dv_ld(fp.cam_pos), d, brute, tr, COUNT ? &c : nullptr);
        }
    }
    append_hit(hit, ray, tr.obj, tr.face, 0.0f, hits, &counters[1]);
    if (COUNT) flush_counters(gstats, c);
}

// fp32 primary ray (direction normalised in fp32); shading recomputes the fp64 ray exactly
__device__ __forceinline__ void primary_dir32(int i, int j, const FrameParams &fp, float &dx, float &dy, float &dz) {
    int hw = fp.W / 2, hh = fp.H / 2;
    float phw = (float)fp.tan_half;
    float phh = phw * (float)fp.H / (float)fp.W;
    float si = phw * ((float)(hw - i) - 0.5f) / (float)hw;
    float sj = phh * ((float)(hh - j) - 0.5f) / (float)hh;
    float x = (float)fp.fwd[0] + (float)fp.left[0] * si + (float)fp.up[0] * sj;
    float y = (float)fp.fwd[1] + (float)fp.left[1] * si + (float)fp.up[1] * sj;
    float z = (float)fp.fwd[2] + (float)fp.left[2] * si + (float)fp.up[2] * sj;
    float inv = rsqrtf(x * x + y * y + z * z);
    dx = x * inv; dy = y * inv; dz = z * inv;
}

// ---------------------------------------------------------------------------------------------
// K2, FP32_FAST, simple variant: one thread per ray (any number of objects)
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128) k_primary_fast_simple(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr,
                                                              HitRec *hits, unsigned int *counters, Counters *gstats) {
    RT_DECLARE_TRAV_SMEM();
    const SceneView sc = bind_frame(sc_in, fr, trav_smem);
    uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nrays = ray_count(fp);
    Counters c; c.nodes = 0; c.tris = 0;
    bool hit = false;
    TraceFast tr; tr.face = -1; tr.obj = -1; tr.t = 0.f;
    if (ray < nrays && !frame_cancelled(counters, fp)) {
        PixMap pm = map_ray(ray, fp, false);
        if (pm.valid && !(fp.tile_w > 0 && block_culled(fp, pm.i - (int)(ray & 7u), pm.j - (int)((ray >> 3) & 3u)))) {
            float dx, dy, dz;
            primary_dir32(pm.i, pm.j, fp, dx, dy, dz);
            dv d = dv_make((double)dx, (double)dy, (double)dz);
            hit = trace_fast_nearest(sc, fp, dv_ld(fp.cam_pos), d, true, tr, COUNT ? &c : nullptr);
        }
    }
    append_hit(hit, ray, tr.obj, tr.face, tr.t, hits, &counters[1]);
    if (COUNT) flush_counters(gstats, c);
}

// ---------------------------------------------------------------------------------------------
// K2, FP32_FAST, persistent warp-cooperative variant (single object / single mesh; RT_FLAG_PERSISTENT).
//
//  * persistent threads: grid = SMs x resident CTAs; every warp pulls rays from one global counter
//  * refill: idle lanes are found with a ballot, ONE atomicAdd per warp claims popc(idle) consecutive rays, the
//    base is shuffled to all lanes and each idle lane takes base + its popc-rank among the idle lanes
//  * rays that miss the mesh's root box (90 % of a frame) leave their lane idle and the warp immediately claims
//    more, so the traversal below always starts with (nearly) full warps
// Round-1 measurements on B200 (profiles/r01_explore_*): a finer-grained version of this kernel (lanes refilled
// while the others are still traversing, postponed-leaf loop with a warp vote, top 256 BFS nodes of the tree in
// shared memory) was 8-15 % SLOWER than the plain one-thread-per-ray kernel above, the shared-memory tree top alone
// cost 0.04 ms, and claiming 256-ray chunks per warp doubled the time (the chip-wide BVH working set stops fitting
// L1/L2).  The hardware block scheduler + L1 already do what the software scheme was meant to; this kernel is kept
// as the selectable north-star design, sharing the per-ray traversal with the default kernel.
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128, 4)
k_primary_fast_persistent(FrameParams fp, MeshView m, float olx, float oly, float olz, HitRec *hits, unsigned int *counters,
                          Counters *gstats) {
    RT_DECLARE_TRAV_SMEM();
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t nrays = ray_count(fp);
    Counters c; c.nodes = 0; c.tris = 0;
    bool have = false;        // lane holds a ray that passed the root box
    uint32_t ray = 0;
    RayF r = make_rayf(0, 0, 0, 0, 0, 1);
    bool exhausted = false;
    for (;;) {
        // ---- refill idle lanes ----
        unsigned idle = __ballot_sync(FULL, !have);
        if (idle && !exhausted) {
            unsigned n = __popc(idle);
            unsigned base = 0;
            if (lane == (unsigned)(__ffs(idle) - 1)) base = atomicAdd(&counters[0], n);
            base = __shfl_sync(FULL, base, __ffs(idle) - 1);
            if (base + n >= nrays) exhausted = true;      // warp-uniform
            if (!have) {
                uint32_t my = base + __popc(idle & lt_mask);
                if (my < nrays) {
                    PixMap pm = map_ray(my, fp, false);
                    if (pm.valid) {
                        float dx, dy, dz;
                        primary_dir32(pm.i, pm.j, fp, dx, dy, dz);
                        r = make_rayf(olx, oly, olz, dx, dy, dz);
                        float tin;
                        // root box; a miss needs no store at all (the framebuffer is pre-cleared)
                        if (slab32(m.lo[0], m.hi[0], m.lo[1], m.hi[1], m.lo[2], m.hi[2], r, 3.0e38f, tin)) { have = true; ray = my; }
                    }
                }
            }
        }
        unsigned live = __ballot_sync(FULL, have);
        if (live == 0) { if (exhausted) break; else continue; }
        if (!exhausted && __popc(live) < 24) continue;    // keep claiming until the warp is (nearly) full
        // ---- traverse: same per-ray loop as the default kernel ----
        float tbest = 3.0e38f;
        int face = -1;
        bool hit = have && mesh_trace_fast<false>(m, r, tbest, face, COUNT ? &c : nullptr, trav_smem);
        append_hit(hit, ray, 0, face, tbest, hits, &counters[1]);
        have = false;
    }
    if (COUNT) flush_counters(gstats, c);
}

// ---------------------------------------------------------------------------------------------
// K4: fused shadow rays + Phong shading, one thread per primary hit; warps claim 32 hits at a time
// ---------------------------------------------------------------------------------------------
template <bool EXACT, bool COUNT, int MINB, bool SHADE32 = false>
__global__ void __launch_bounds__(128, MINB) k_shade(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr, bool brute,
                                                      const HitRec *__restrict__ hits, unsigned int *counters,
                                                      double *__restrict__ hit_colour, uint32_t *__restrict__ hit_lit,
                                                      Counters *gstats) {
    RT_DECLARE_TRAV_SMEM();
    const SceneView sc = bind_frame(sc_in, fr, EXACT ? nullptr : trav_smem);
    const unsigned lane = threadIdx.x & 31;
    const unsigned int nhits = counters[1];
    Counters c; c.nodes = 0; c.tris = 0;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = frame_cancelled(counters, fp) ? 0xffffffffu : atomicAdd(&counters[2], 32u);
        base = __shfl_sync(FULL, base, 0);
        if (base >= nhits) break;
        // every lane runs shade_hit (its light loop reconverges with __syncwarp); lanes past the end redo the last hit
        const bool valid = base + lane < nhits;
        unsigned h = valid ? base + lane : nhits - 1;
        HitRec hr = hits[h];
        PixMap pm = map_ray(hr.pix, fp, false);
        double col[3];
        uint32_t lit = SHADE32 ? shade_hit_fast32(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, col, (COUNT && valid) ? &c : nullptr)
                               : shade_hit<EXACT>(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, brute, col, (COUNT && valid) ? &c : nullptr);
        if (!EXACT) tune_any_defer(sc.tune, sc.tune_host, valid, lit, fp.n_lights);
        if (valid) {
            hit_colour[3 * (size_t)h] = col[0]; hit_colour[3 * (size_t)h + 1] = col[1]; hit_colour[3 * (size_t)h + 2] = col[2];
            hit_lit[h] = lit;
        }
    }
    if (COUNT) flush_counters(gstats ? gstats + 1 : nullptr, c);
}

// ---------------------------------------------------------------------------------------------
// K4, split form (RT_B200_SPLIT_SHADE=1): the same arithmetic as k_shade in three kernels, so that the fp64 parts run with
// fully converged warps and the shadow rays get a lean, high-occupancy traversal kernel with one thread per (hit, light):
//   k_hit_geometry : per hit   -- fp64 re-intersection of the winner, normal                      -> hit_geom (P, N)
//   k_shadow       : per (hit, light), light-major so that a warp holds 32 neighbouring hits of ONE light -> lit bits
//   k_phong        : per hit   -- Phong accumulation over the lights in order, using the lit bits  -> hit colour
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_hit_geometry(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr,
                                                       const HitRec *__restrict__ hits, const unsigned int *counters,
                                                       double *__restrict__ hit_geom, uint32_t *__restrict__ hit_lit) {
    const SceneView sc = bind_frame(sc_in, fr);
    const unsigned int nhits = counters[1];
    for (unsigned h = blockIdx.x * blockDim.x + threadIdx.x; h < nhits; h += gridDim.x * blockDim.x) {
        HitRec hr = hits[h];
        PixMap pm = map_ray(hr.pix, fp, false);
        HitGeom g = hit_geometry(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t);
        double *q = hit_geom + 6 * (size_t)h;
        q[0] = g.P.x; q[1] = g.P.y; q[2] = g.P.z; q[3] = g.N.x; q[4] = g.N.y; q[5] = g.N.z;
        hit_lit[h] = 0u;
    }
}

template <bool EXACT, bool COUNT>
__global__ void __launch_bounds__(128) k_shadow(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr, bool brute,
                                                 const HitRec *__restrict__ hits, unsigned int *counters,
                                                 const double *__restrict__ hit_geom, uint32_t *hit_lit, Counters *gstats) {
    const SceneView sc = bind_frame(sc_in, fr);
    const unsigned lane = threadIdx.x & 31;
    const unsigned int nhits = counters[1];
    const unsigned long long npairs = (unsigned long long)nhits * (unsigned long long)fp.n_lights;
    Counters c; c.nodes = 0; c.tris = 0;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&counters[2], 32u);
        base = __shfl_sync(FULL, base, 0);
        if (base >= npairs) break;
        unsigned long long q = (unsigned long long)base + lane;
        if (q < npairs) {
            const int l = (int)(q / nhits);
            const unsigned h = (unsigned)(q % nhits);
            const double *g = hit_geom + 6 * (size_t)h;
            dv Ld;
            if (light_visible<EXACT>(sc, fp, dv_make(g[0], g[1], g[2]), hits[h].obj, l, brute, Ld, COUNT ? &c : nullptr))
                atomicOr(&hit_lit[h], 1u << l);
        }
    }
    if (COUNT) flush_counters(gstats ? gstats + 1 : nullptr, c);
}

__global__ void __launch_bounds__(128) k_phong(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr,
                                                const HitRec *__restrict__ hits, const unsigned int *counters,
                                                const double *__restrict__ hit_geom, const uint32_t *__restrict__ hit_lit,
                                                double *__restrict__ hit_colour) {
    const SceneView sc = bind_frame(sc_in, fr);
    const unsigned int nhits = counters[1];
    for (unsigned h = blockIdx.x * blockDim.x + threadIdx.x; h < nhits; h += gridDim.x * blockDim.x) {
        HitRec hr = hits[h];
        const double *q = hit_geom + 6 * (size_t)h;
        HitGeom g;
        g.P = dv_make(q[0], q[1], q[2]); g.N = dv_make(q[3], q[4], q[5]);
        double col[3];
        phong_from_mask(sc, fp, g, hit_material(sc, hr.obj, hr.face), hit_lit[h], col);
        hit_colour[3 * (size_t)h] = col[0]; hit_colour[3 * (size_t)h + 1] = col[1]; hit_colour[3 * (size_t)h + 2] = col[2];
    }
}

// ---------------------------------------------------------------------------------------------
// Fused frame kernel (FP32_FAST; RT_FLAG_FUSED / RT_B200_FUSED): K2 + K4 + K5 in ONE launch.  A warp owns an 8x4 pixel
// block: it traces the 32 primary rays, and if any lane hit something, the hit lanes shade their pixel (fp64
// re-intersection, one shadow ray per light, Phong; lanes without a hit shadow the first hit lane so that the warp stays
// converged through shade_hit's per-light __syncwarp) and write the three bytes of their pixel.  No hit list, no colour
// buffer, no second and third launch: a frame is one kernel, so its latency is one kernel tail instead of three -- what a
// rank of an 8-GPU box needs when its share of a 4K frame is 0.1 ms of work.  Blocks outside the projected
// root boxes are never launched (region mode) or exit at once (tile mode).
// ---------------------------------------------------------------------------------------------
#ifndef RT_FUSED_THREADS
#define RT_FUSED_THREADS 128      // threads per CTA of the fused kernel (64: measured in profiles/r02_kernel_ab.txt)
#endif
// one 8x4 pixel block of a frame, all 32 lanes of a warp: `ray` = 32 * block + lane in the numbering of map_ray
template <bool COUNT, bool DEFER>
__device__ __forceinline__ void fused_block(const FrameParams &fp, const SceneView &sc, uint32_t ray, bool rowmajor, bool sdl, unsigned int *counters,
                                            uint8_t *__restrict__ out_rgb, double *__restrict__ out_colour, int32_t *__restrict__ out_face,
                                            int32_t *__restrict__ out_obj, uint32_t *__restrict__ out_lit, Counters &c0, Counters &c1) {
    const unsigned lane = threadIdx.x & 31;
    const PixMap pm = map_ray(ray, fp, rowmajor);
    if (fp.tile_w > 0 && block_culled(fp, pm.i - (int)(ray & 7u), pm.j - (int)((ray >> 3) & 3u))) return;      // warp-uniform: a warp is one 8x4 block
    bool hit = false;
    TraceFast tr; tr.face = -1; tr.obj = -1; tr.t = 0.f;
    if (pm.valid) {
        float dx, dy, dz;
        primary_dir32(pm.i, pm.j, fp, dx, dy, dz);
        hit = trace_fast_nearest(sc, fp, dv_ld(fp.cam_pos), dv_make((double)dx, (double)dy, (double)dz), true, tr, COUNT ? &c0 : nullptr);
    }
    const unsigned m = __ballot_sync(FULL, hit);
    if (m != 0) {
        if (lane == 0) atomicAdd(&counters[1], (unsigned)__popc(m));          // hit pixels of the frame (rt_stats, ray accounting)
        const int src = hit ? (int)lane : __ffs(m) - 1;
        const int si = __shfl_sync(FULL, pm.i, src), sj = __shfl_sync(FULL, pm.j, src);
        const int sobj = __shfl_sync(FULL, tr.obj, src), sface = __shfl_sync(FULL, tr.face, src);
        const float st = __shfl_sync(FULL, tr.t, src);
        double col[3];
        const uint32_t lit = shade_hit<false, DEFER>(sc, fp, si, sj, sobj, sface, st, false, col, (COUNT && hit) ? &c1 : nullptr);
        tune_any_defer(sc.tune, sc.tune_host, hit, lit, fp.n_lights);
        if (hit) {
            const size_t o = pm.out;
            if (out_rgb) {
                out_rgb[3 * o] = quantise_exact(col[0], sdl);
                out_rgb[3 * o + 1] = quantise_exact(col[1], sdl);
                out_rgb[3 * o + 2] = quantise_exact(col[2], sdl);
            }
            if (out_colour) { out_colour[3 * o] = col[0]; out_colour[3 * o + 1] = col[1]; out_colour[3 * o + 2] = col[2]; }
            if (out_face) out_face[o] = tr.face;
            if (out_obj) out_obj[o] = tr.obj;
            if (out_lit) out_lit[o] = lit;
        }
    }
}

template <bool COUNT, int MINB, bool DEFER>
__global__ void __launch_bounds__(RT_FUSED_THREADS, MINB * (128 / RT_FUSED_THREADS)) k_frame_fused(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr,
                                                            bool rowmajor, bool sdl, unsigned int *counters,
                                                            uint8_t *__restrict__ out_rgb, double *__restrict__ out_colour,
                                                            int32_t *__restrict__ out_face, int32_t *__restrict__ out_obj,
                                                            uint32_t *__restrict__ out_lit, Counters *gstats) {
    RT_DECLARE_TRAV_SMEM();
    const SceneView sc = bind_frame(sc_in, fr, trav_smem);
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nrays = ray_count(fp);
    if ((ray & ~31u) >= nrays) return;                        // warp-uniform
    if (warp_cancelled(counters, fp)) return;
    Counters c0, c1; c0.nodes = c0.tris = c1.nodes = c1.tris = 0;
    fused_block<COUNT, DEFER>(fp, sc, ray, rowmajor, sdl, counters, out_rgb, out_colour, out_face, out_obj, out_lit, c0, c1);
    if (COUNT) { flush_counters(gstats, c0); flush_counters(gstats ? gstats + 1 : nullptr, c1); }
}

// ---------------------------------------------------------------------------------------------
// The same frame with PERSISTENT, SM-AFFINE warps (RT_FLAG_SM_AFFINE / RT_B200_SM_AFFINE; north_star part 2: persistent threads
// that refill their work with one atomic per warp).  The hardware block scheduler hands the CTAs of the plain kernel to whichever
// SM has room, so the 36 warps resident on one SM work on ~9 unrelated places of the screen and every SM pulls the same BVH
// lines through its own L1.  Here the grid is SMs x resident CTAs; the frame is cut into CHUNKS of RT_CHUNK_BLOCKS 8x4 blocks that
// are neighbours on the screen (tile mode: consecutive blocks of the rank's tiles; region mode: 4 x RT_CHUNK_BLOCKS/4 blocks = 32 x 32 pixels),
// every SM owns one chunk at a time (its id sits in a per-SM word of the lane's chunk state) and ALL its warps take their next
// block from that chunk with one atomicAdd -- so the warps that share an L1 walk the same part of the tree at the same time.
// When a chunk runs out, the warp that notices locks the word, takes the next chunk of the frame from a global counter and
// publishes it; when the frame has no chunks left a warp helps the other SMs finish theirs before it retires.
// chunk state (uint32 words, zeroed per frame): [0] next chunk, [1] spare, [2 .. 2+RT_MAX_SM) per-SM word, then one
// next-block counter per chunk.  Per-SM word: 0 = nothing yet, 1 = locked (being refilled), 2 = frame exhausted, c + 3 = chunk c.
// ---------------------------------------------------------------------------------------------
struct ChunkGrid { uint32_t n_chunks, n_blocks, nbx, nby, cgx; };     // region mode: cgx chunks per row of chunks; tile mode: cgx = 0

__host__ __device__ inline ChunkGrid chunk_grid(const FrameParams &fp) {
    ChunkGrid g;
    g.n_blocks = ray_count(fp) / 32u;
    if (fp.tile_w > 0) { g.nbx = g.nby = g.cgx = 0; g.n_chunks = (g.n_blocks + RT_CHUNK_BLOCKS - 1) / RT_CHUNK_BLOCKS; return g; }
    g.nbx = fp.rnbx > 0 ? (uint32_t)fp.rnbx : (uint32_t)(fp.w + 7) / 8;
    g.nby = g.nbx ? g.n_blocks / g.nbx : 0;
    g.cgx = (g.nbx + 3) / 4;
    g.n_chunks = g.cgx * ((g.nby + RT_CHUNK_BLOCKS / 4 - 1) / (RT_CHUNK_BLOCKS / 4));
    return g;
}
// block `idx` of chunk `c` -> linear block number of map_ray, or 0xffffffff if that slot lies outside the frame
__device__ __forceinline__ uint32_t chunk_block(const ChunkGrid &g, uint32_t c, uint32_t idx) {
    if (g.cgx == 0) { const uint32_t b = c * RT_CHUNK_BLOCKS + idx; return b < g.n_blocks ? b : 0xffffffffu; }
    const uint32_t bx = (c % g.cgx) * 4 + (idx & 3u), by = (c / g.cgx) * (RT_CHUNK_BLOCKS / 4) + (idx >> 2);
    return (bx < g.nbx && by < g.nby) ? by * g.nbx + bx : 0xffffffffu;
}
__device__ __forceinline__ unsigned int ld_volatile(const unsigned int *p) { return *(const volatile unsigned int *)p; }

// lane 0 of a warp: the next block of this SM's chunk (0xfffffffe = the frame is finished).  `slots` = blocks per chunk.
__device__ __forceinline__ uint32_t claim_block(unsigned int *state, const ChunkGrid &g, unsigned smid, unsigned n_sm) {
    unsigned int *word = state + 2 + smid, *next = state + 2 + RT_MAX_SM;
    for (;;) {
        const unsigned int w = ld_volatile(word);
        if (w >= 3u) {
            const uint32_t c = w - 3u, idx = atomicAdd(&next[c], 1u);
            if (idx < RT_CHUNK_BLOCKS) { const uint32_t b = chunk_block(g, c, idx); if (b != 0xffffffffu) return b; continue; }
            if (atomicCAS(word, w, 1u) != w) continue;                         // somebody else refills
        } else if (w == 1u) {
            __nanosleep(40);
            continue;
        } else if (w == 2u) {
            break;
        } else if (atomicCAS(word, 0u, 1u) != 0u) continue;
        // this warp holds the lock: take the next chunk of the frame
        const uint32_t c = atomicAdd(&state[0], 1u);
        __threadfence();
        *(volatile unsigned int *)word = c < g.n_chunks ? c + 3u : 2u;
    }
    // no chunks left: help the SMs that are still inside one
    for (unsigned k = 1; k < n_sm; k++) {
        const unsigned int w = ld_volatile(state + 2 + (smid + k) % n_sm);
        if (w < 3u) continue;
        const uint32_t c = w - 3u;
        while (ld_volatile(&next[c]) < RT_CHUNK_BLOCKS) {
            const uint32_t idx = atomicAdd(&next[c], 1u);
            if (idx >= RT_CHUNK_BLOCKS) break;
            const uint32_t b = chunk_block(g, c, idx);
            if (b != 0xffffffffu) return b;
        }
    }
    return 0xfffffffeu;
}

template <bool COUNT, int MINB>
__global__ void __launch_bounds__(128, MINB) k_frame_fused_sm(FrameParams fp, SceneView sc_in, const __grid_constant__ InlineFrame fr, bool rowmajor, bool sdl,
                                                               unsigned int *counters, unsigned int *chunk_state, unsigned n_sm,
                                                               uint8_t *__restrict__ out_rgb, double *__restrict__ out_colour,
                                                               int32_t *__restrict__ out_face, int32_t *__restrict__ out_obj,
                                                               uint32_t *__restrict__ out_lit, Counters *gstats) {
    const SceneView sc = bind_frame(sc_in, fr, nullptr);
    const ChunkGrid g = chunk_grid(fp);
    const unsigned lane = threadIdx.x & 31;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    Counters c0, c1; c0.nodes = c0.tris = c1.nodes = c1.tris = 0;
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = frame_cancelled(counters, fp) ? 0xfffffffeu : claim_block(chunk_state, g, smid % RT_MAX_SM, n_sm);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk == 0xfffffffeu) break;
        fused_block<COUNT, true>(fp, sc, blk * 32u + lane, rowmajor, sdl, counters, out_rgb, out_colour, out_face, out_obj, out_lit, c0, c1);
        __syncwarp();
    }
    if (COUNT) { flush_counters(gstats, c0); flush_counters(gstats ? gstats + 1 : nullptr, c1); }
}

// ---------------------------------------------------------------------------------------------
// K5: 8-bit framebuffer writer.  Misses were pre-cleared (sequential/main.go:17 FillRect(nil,0);
// distributed/main.go:70 default 0,0,0).  Quantisers: colour.go:59-61 and colour.go:54-56 (+ SDL >>8).
// ---------------------------------------------------------------------------------------------
__global__ void k_write(FrameParams fp, const HitRec *__restrict__ hits, const unsigned int *counters,
                        const double *__restrict__ hit_colour, const uint32_t *__restrict__ hit_lit, bool rowmajor, bool sdl,
                        uint8_t *__restrict__ out_rgb, double *__restrict__ out_colour, int32_t *__restrict__ out_face,
                        int32_t *__restrict__ out_obj, uint32_t *__restrict__ out_lit) {
    const unsigned int nhits = counters[1];
    for (unsigned h = blockIdx.x * blockDim.x + threadIdx.x; h < nhits; h += gridDim.x * blockDim.x) {
        HitRec hr = hits[h];
        PixMap pm = map_ray(hr.pix, fp, rowmajor);
        size_t o = pm.out;
        double cr = hit_colour[3 * (size_t)h], cg = hit_colour[3 * (size_t)h + 1], cb = hit_colour[3 * (size_t)h + 2];
        if (out_rgb) {
            out_rgb[3 * o] = quantise_exact(cr, sdl);
            out_rgb[3 * o + 1] = quantise_exact(cg, sdl);
            out_rgb[3 * o + 2] = quantise_exact(cb, sdl);
        }
        if (out_colour) { out_colour[3 * o] = cr; out_colour[3 * o + 1] = cg; out_colour[3 * o + 2] = cb; }
        if (out_face) out_face[o] = hr.face;
        if (out_obj) out_obj[o] = hr.obj;
        if (out_lit) out_lit[o] = hit_lit[h];
    }
}

// multi-GPU: gathered interleaved-tile buffers -> framebuffer.  One CTA per tile: the packed tile (tw*th*3 bytes,
// contiguous) is staged in shared memory with 16-byte loads, then written out as 32-bit words that are contiguous in
// the destination layout (a pixel column of th*3 bytes for the column-major BulkTrace layout, a pixel row of tw*3
// bytes for the row-major one).  Tiles cut by the screen edge take the byte-wise path.
__global__ void __launch_bounds__(128) k_untile(const uint8_t *__restrict__ tiles, uint8_t *__restrict__ out, int W, int H, int tw,
                                                 int th, int n_ranks, int tiles_per_rank_max, bool rowmajor) {
    extern __shared__ __align__(16) uint8_t s_tile[];
    const int tiles_x = (W + tw - 1) / tw;
    const int TX = blockIdx.x % tiles_x, TY = blockIdx.x / tiles_x;      // one CTA per tile POSITION
    const int t = (int)tile_id((uint32_t)TX, (uint32_t)TY, (uint32_t)tiles_x);
    const int tx0 = TX * tw, ty0 = TY * th;
    const int rank = t % n_ranks, lt = t / n_ranks;
    const size_t tile_bytes = (size_t)tw * th * 3;
    const uint8_t *src = tiles + ((size_t)rank * tiles_per_rank_max + lt) * tile_bytes;
    const uint4 *src4 = reinterpret_cast<const uint4 *>(src);           // tile_bytes is a multiple of 96
    uint4 *dst4 = reinterpret_cast<uint4 *>(s_tile);
    for (int k = threadIdx.x; k < (int)(tile_bytes / 16); k += blockDim.x) dst4[k] = __ldg(src4 + k);
    __syncthreads();
    const bool interior = tx0 + tw <= W && ty0 + th <= H && (H % 4 == 0) && (W % 4 == 0);
    if (interior) {
        if (rowmajor) {
            const int words_per_row = tw * 3 / 4;                        // tw is a multiple of 8
            for (int k = threadIdx.x; k < th * words_per_row; k += blockDim.x) {
                int row = k / words_per_row, wd = k % words_per_row;
                uint32_t v = *reinterpret_cast<const uint32_t *>(s_tile + (size_t)row * tw * 3 + 4 * wd);
                *reinterpret_cast<uint32_t *>(out + ((size_t)(ty0 + row) * W + tx0) * 3 + 4 * wd) = v;
            }
        } else {
            const int words_per_col = th * 3 / 4;                        // th is a multiple of 4
            for (int k = threadIdx.x; k < tw * words_per_col; k += blockDim.x) {
                int col = k / words_per_col, wd = k % words_per_col;
                uint32_t v = 0;
#pragma unroll
                for (int bb = 0; bb < 4; bb++) {
                    int byte = 4 * wd + bb;                              // byte within the column: pixel byte/3, channel byte%3
                    int row = byte / 3, ch = byte % 3;
                    v |= (uint32_t)s_tile[((size_t)row * tw + col) * 3 + ch] << (8 * bb);
                }
                *reinterpret_cast<uint32_t *>(out + ((size_t)(tx0 + col) * H + ty0) * 3 + 4 * wd) = v;
            }
        }
    } else {
        for (int k = threadIdx.x; k < tw * th; k += blockDim.x) {
            int col = k % tw, row = k / tw;
            int i = tx0 + col, j = ty0 + row;
            if (i >= W || j >= H) continue;
            size_t dst = (rowmajor ? (size_t)j * W + i : (size_t)i * H + j) * 3;
            const uint8_t *p = s_tile + (size_t)k * 3;
            out[dst] = p[0]; out[dst + 1] = p[1]; out[dst + 2] = p[2];
        }
    }
}

// multi-GPU, peer-memory push: copies THIS rank's packed tiles (local) into the full framebuffer `out` in its final
// layout; `out` normally lives on rank 0's GPU (CUDA IPC mapping), so the stores travel over NVLink.  One CTA per local
// tile: staged in shared memory with 16-byte loads, written as 16-byte vectors that exactly cover the tile's column
// (column-major BulkTrace layout) or row (row-major) segments -- wide, aligned, coalesced transactions are what the link
// wants (scattered 1-byte stores from the framebuffer writer cost 17 % of the frame at N = 2, profiles/r01_peer_*).
template <typename VEC>
__device__ __forceinline__ void push_segments(const uint8_t *s_tile, uint8_t *out, int n_seg, int seg_bytes, size_t seg_stride_out,
                                              size_t out_base, int tw, bool rowmajor) {
    const int vpb = (int)sizeof(VEC);
    const int vecs_per_seg = seg_bytes / vpb;
    for (int k = threadIdx.x; k < n_seg * vecs_per_seg; k += blockDim.x) {
        const int seg = k / vecs_per_seg, v = k % vecs_per_seg;
        VEC val;
        uint8_t *vb = reinterpret_cast<uint8_t *>(&val);
        if (rowmajor) {
            val = *reinterpret_cast<const VEC *>(s_tile + (size_t)seg * seg_bytes + (size_t)v * vpb);      // a tile row is contiguous
        } else {
#pragma unroll
            for (int b = 0; b < vpb; b++) {
                const int byte = v * vpb + b;                       // byte within the column segment
                vb[b] = s_tile[((size_t)(byte / 3) * tw + seg) * 3 + byte % 3];
            }
        }
        *reinterpret_cast<VEC *>(out + out_base + (size_t)seg * seg_stride_out + (size_t)v * vpb) = val;
    }
}

__device__ __forceinline__ void push_one_tile(const FrameParams &fp, bool rowmajor, const uint8_t *__restrict__ tiles,
                                              uint8_t *__restrict__ out, uint8_t *s_tile, int lt) {
    const uint32_t t = (uint32_t)lt * (uint32_t)fp.n_ranks + (uint32_t)fp.rank;
    const int tw = fp.tile_w, th = fp.tile_h, W = fp.w, H = fp.h;      // the framebuffer is the region's (full frame: the screen)
    uint32_t TX, TY;
    tile_pos(t, (uint32_t)fp.tiles_x, TX, TY);
    const int tx0 = (int)TX * tw, ty0 = (int)TY * th;
    if (tx0 >= W || ty0 >= H) return;
    const size_t tile_bytes = (size_t)tw * th * 3;
    const uint4 *src4 = reinterpret_cast<const uint4 *>(tiles + (size_t)lt * tile_bytes);
    uint4 *dst4 = reinterpret_cast<uint4 *>(s_tile);
    for (int k = threadIdx.x; k < (int)(tile_bytes / 16); k += blockDim.x) dst4[k] = __ldg(src4 + k);
    __syncthreads();
    const bool interior = tx0 + tw <= W && ty0 + th <= H;
    const int n_seg = rowmajor ? th : tw, seg_bytes = (rowmajor ? tw : th) * 3;
    const size_t stride = (size_t)(rowmajor ? W : H) * 3;
    const size_t base = (rowmajor ? (size_t)ty0 * W + tx0 : (size_t)tx0 * H + ty0) * 3;
    if (interior && seg_bytes % 16 == 0 && stride % 16 == 0 && base % 16 == 0)
        push_segments<uint4>(s_tile, out, n_seg, seg_bytes, stride, base, tw, rowmajor);
    else if (interior && stride % 4 == 0 && base % 4 == 0)
        push_segments<uint32_t>(s_tile, out, n_seg, seg_bytes, stride, base, tw, rowmajor);
    else
        for (int k = threadIdx.x; k < tw * th; k += blockDim.x) {
            int i = tx0 + k % tw, j = ty0 + k / tw;
            if (i >= W || j >= H) continue;
            size_t dst = (rowmajor ? (size_t)j * W + i : (size_t)i * H + j) * 3;
            const uint8_t *p = s_tile + (size_t)k * 3;
            out[dst] = p[0]; out[dst + 1] = p[1]; out[dst + 2] = p[2];
        }
}

// ---- frame hand-off flags in peer memory -------------------------------------------------------------------------
// One 256-byte record per framebuffer slot, living next to the framebuffers on the gathering rank (CUDA IPC, reached over
// NVLink by the others):  uint32 arrive[32] (arrive[r] = number of the last frame rank r has completely pushed into the
// slot), consumed (number of the last frame the gathering rank is done with: the slot may be overwritten), error.
// Frame numbers count per slot from 1.  Release/acquire at system scope orders the tile stores before the flag.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// waits until *p has reached `want` (wrap-safe); gives up after timeout_ms (a dead peer must not hang this GPU:
// the reference's own order time-out is 2 s, master/main.go:35)
__device__ bool spin_until(const uint32_t *p, uint32_t want, uint32_t timeout_ms) {
    if ((int32_t)(ld_acquire_sys(p) - want) >= 0) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        if ((int32_t)(ld_acquire_sys(p) - want) >= 0) return true;
        __nanosleep(256);
        if (globaltimer_ns() - t0 > (unsigned long long)timeout_ms * 1000000ull) return false;
    }
}

// rank 0 (the gathering rank): frame `seq` of the slot is complete once every rank's arrive flag has reached it;
// with `release` the slot is handed back at once (device-only consumer), else rt_frame_release_device does it after the consumer.
__global__ void k_frame_complete(uint32_t *slot, uint32_t seq, int n_ranks, int release, uint32_t timeout_ms) {
    const int r = (int)threadIdx.x;
    if (r < n_ranks && !spin_until(slot + RT_SYNC_ARRIVE + r, seq, timeout_ms))
        *(volatile uint32_t *)(slot + RT_SYNC_ERROR) = 0x100u + (uint32_t)r;
    __syncwarp();
    if (release && r == 0) { __threadfence_system(); st_release_sys(slot + RT_SYNC_CONSUMED, seq); }
}
__global__ void k_frame_release(uint32_t *slot, uint32_t seq) { __threadfence_system(); st_release_sys(slot + RT_SYNC_CONSUMED, seq); }
// the two halves of the hand-off as stand-alone kernels, for a hand-over that is not a kernel (copy-engine transfer into
// the host framebuffer): wait for "previous frame of the slot consumed" before it, publish "rank arrived" after it
__global__ void k_wait_consumed(uint32_t *slot, uint32_t seq, uint32_t rank, uint32_t timeout_ms) {
    if (!spin_until(slot + RT_SYNC_CONSUMED, seq - 1u, timeout_ms)) *(volatile uint32_t *)(slot + RT_SYNC_ERROR) = 0x200u + rank;
}
__global__ void k_signal_arrive(uint32_t *slot, uint32_t seq, uint32_t rank) { __threadfence_system(); st_release_sys(slot + RT_SYNC_ARRIVE + rank, seq); }

__global__ void __launch_bounds__(128) k_push_tiles(FrameParams fp, bool rowmajor, const uint8_t *__restrict__ tiles,
                                                     uint8_t *__restrict__ out, PushSync sy) {
    extern __shared__ __align__(16) uint8_t s_tile[];
    if (sy.slot && !(sy.debug & 1u)) {              // the slot's previous frame must have been consumed before it is overwritten
        if (threadIdx.x == 0 && !spin_until(sy.slot + RT_SYNC_CONSUMED, sy.seq - 1u, sy.timeout_ms))
            *(volatile uint32_t *)(sy.slot + RT_SYNC_ERROR) = 0x200u + (uint32_t)fp.rank;
        __syncthreads();
    }
    // with a hand-off the grid is a few CTAs per SM looping over the tiles: the remote flag poll above and the system-scope
    // fence below (each a NVLink round trip) are paid per CTA, not per tile (per tile they cost 6 % of the N = 2 frame)
    for (int lt = blockIdx.x; lt < fp.n_local_tiles; lt += gridDim.x) {
        push_one_tile(fp, rowmajor, tiles, out, s_tile, lt);
        __syncthreads();                            // s_tile is restaged by the next tile
    }
    if (sy.slot) {                                  // last CTA to finish publishes "rank has pushed frame seq"
        if (sy.debug & 2u) __threadfence(); else __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(sy.blocks_done, 1u);
            if (prev == gridDim.x - 1) {
                *sy.blocks_done = 0u;               // next launch on this lane is stream-ordered after this one
                __threadfence_system();
                st_release_sys(sy.slot + RT_SYNC_ARRIVE + fp.rank, sy.seq);
            }
        }
    }
}
// multi-GPU, peer-memory writer: clears the pixels of THIS rank's tiles in the full framebuffer `out` (which may live on
// another GPU, mapped through CUDA IPC / NVLink).  One CTA per local tile; interior tiles are cleared with 32-bit words
// that exactly cover the tile's column (or row) segments, so no byte owned by another rank is ever touched.
__global__ void __launch_bounds__(128) k_clear_tiles(FrameParams fp, bool rowmajor, uint8_t *__restrict__ out) {
    const int lt = blockIdx.x;
    const uint32_t t = (uint32_t)lt * (uint32_t)fp.n_ranks + (uint32_t)fp.rank;
    const int tw = fp.tile_w, th = fp.tile_h, W = fp.w, H = fp.h;      // region-relative, like map_ray
    uint32_t TX, TY;
    tile_pos(t, (uint32_t)fp.tiles_x, TX, TY);
    const int tx0 = (int)TX * tw, ty0 = (int)TY * th;
    if (tx0 >= W || ty0 >= H) return;
    const bool interior = tx0 + tw <= W && ty0 + th <= H && (H % 4 == 0) && (W % 4 == 0);
    if (interior) {
        if (rowmajor) {
            const int wpr = tw * 3 / 4;
            for (int k = threadIdx.x; k < th * wpr; k += blockDim.x)
                *reinterpret_cast<uint32_t *>(out + ((size_t)(ty0 + k / wpr) * W + tx0) * 3 + 4 * (k % wpr)) = 0u;
        } else {
            const int wpc = th * 3 / 4;
            for (int k = threadIdx.x; k < tw * wpc; k += blockDim.x)
                *reinterpret_cast<uint32_t *>(out + ((size_t)(tx0 + k / wpc) * H + ty0) * 3 + 4 * (k % wpc)) = 0u;
        }
    } else {
        for (int k = threadIdx.x; k < tw * th; k += blockDim.x) {
            int i = tx0 + k % tw, j = ty0 + k / tw;
            if (i >= W || j >= H) continue;
            size_t dst = (rowmajor ? (size_t)j * W + i : (size_t)i * H + j) * 3;
            out[dst] = 0; out[dst + 1] = 0; out[dst + 2] = 0;
        }
    }
}

// FMA-pipe peak: 8 independent dependent-chains per thread, all operands in registers
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T *out, int iters, T a, T b) {
    T x0 = a + (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3, x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6, x7 = x0 + (T)7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)123456789) out[0] = s;   // never true: keeps the chain alive
}

}  // namespace

cudaError_t rt_fma_peak(bool fp64, int sm_count, double *tflops) {
    const int iters = 1 << 14, threads = 256, blocks = sm_count * 8;
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        if (fp64) k_fma_peak<double><<<blocks, threads>>>((double *)d, iters, 0.999999, 1e-9);
        else k_fma_peak<float><<<blocks, threads>>>((float *)d, iters, 0.999999f, 1e-9f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    e = cudaGetLastError();
    *tflops = 2.0 * 8.0 * (double)iters * threads * blocks / (best * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return e;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <class K> static void prefer_l1(K kern) {
    static thread_local int done_for_device = -1;             // per kernel instantiation (K differs) and calling thread
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev == done_for_device) return;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    cudaGetLastError();
    done_for_device = dev;
}

cudaError_t rt_launch_primary(const RenderJob &job, cudaStream_t st) {
    const FrameParams &fp = job.fp;
    SceneView sc; sc.meshes = job.d_meshes; sc.objs = job.d_objs; sc.lights = job.d_lights; sc.trav_smem = nullptr; sc.tune = job.d_tune; sc.tune_host = job.d_tune_host; sc.defer = 0;
    const uint32_t nrays = ray_count(fp);
    if (nrays == 0) return cudaSuccess;
    const bool count = (job.flags & RT_FLAG_COUNT) != 0;
    const unsigned nblk = (nrays + 127) / 128;
    if (job.precision == RT_FP64_EXACT) {
        const bool brute = (job.flags & RT_FLAG_BRUTE_FORCE) != 0;
        if (count) k_primary_exact<true><<<nblk, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_stats);
        else k_primary_exact<false><<<nblk, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_stats);
        return cudaGetLastError();
    }
    const bool want_persistent = (job.flags & RT_FLAG_PERSISTENT) != 0 || (job.primary_persistent && !(job.flags & RT_FLAG_SIMPLE));
    const bool simple = !want_persistent || fp.n_obj != 1 || !job.single_mesh;
    if (simple) {
        if (count) k_primary_fast_simple<true><<<nblk, 128, 0, st>>>(fp, sc, job.fr, job.d_hits, job.d_counters, job.d_stats);
        else k_primary_fast_simple<false><<<nblk, 128, 0, st>>>(fp, sc, job.fr, job.d_hits, job.d_counters, job.d_stats);
        return cudaGetLastError();
    }
    // persistent: one object; its mesh view and the object-space ray origin are passed by value
    const MeshView &m = job.h_mesh0;
    const float olx = (float)(fp.cam_pos[0] - job.h_obj0_pos[0]), oly = (float)(fp.cam_pos[1] - job.h_obj0_pos[1]),
                olz = (float)(fp.cam_pos[2] - job.h_obj0_pos[2]);
    int per_sm = 0;
    cudaError_t e;
#define LAUNCH_PERSISTENT(COUNTF)                                                                                        \
    do {                                                                                                                 \
        auto kern = k_primary_fast_persistent<COUNTF>;                                                                   \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0);                                        \
        if (e != cudaSuccess) return e;                                                                                  \
        if (per_sm < 1) per_sm = 1;                                                                                      \
        unsigned g = (unsigned)(job.sm_count * per_sm);                                                                  \
        unsigned need = (nrays + 127) / 128;                                                                             \
        if (g > need) g = need;                                                                                          \
        kern<<<g, 128, 0, st>>>(fp, m, olx, oly, olz, job.d_hits, job.d_counters, job.d_stats);                          \
    } while (0)
    if (count) LAUNCH_PERSISTENT(true); else LAUNCH_PERSISTENT(false);
#undef LAUNCH_PERSISTENT
    return cudaGetLastError();
}

cudaError_t rt_launch_shade(const RenderJob &job, cudaStream_t st) {
    const FrameParams &fp = job.fp;
    SceneView sc; sc.meshes = job.d_meshes; sc.objs = job.d_objs; sc.lights = job.d_lights; sc.trav_smem = nullptr; sc.tune = job.d_tune; sc.tune_host = job.d_tune_host; sc.defer = 0;
    if (ray_count(fp) == 0) return cudaSuccess;
    const bool count = (job.flags & RT_FLAG_COUNT) != 0;
    const bool brute = (job.flags & RT_FLAG_BRUTE_FORCE) != 0;
    const unsigned g = (unsigned)job.sm_count * 8u;
    if (job.split_shade && fp.n_lights <= 32 && job.d_hit_geom) {
        k_hit_geometry<<<g, 128, 0, st>>>(fp, sc, job.fr, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit);
        const unsigned gs = (unsigned)job.sm_count * 12u;
        if (job.precision == RT_FP64_EXACT) {
            if (count) k_shadow<true, true><<<gs, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit, job.d_stats);
            else k_shadow<true, false><<<gs, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit, job.d_stats);
        } else {
            if (count) k_shadow<false, true><<<gs, 128, 0, st>>>(fp, sc, job.fr, false, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit, job.d_stats);
            else k_shadow<false, false><<<gs, 128, 0, st>>>(fp, sc, job.fr, false, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit, job.d_stats);
        }
        k_phong<<<g, 128, 0, st>>>(fp, sc, job.fr, job.d_hits, job.d_counters, job.d_hit_geom, job.d_hit_lit, job.d_hit_colour);
        return cudaGetLastError();
    }
#define LAUNCH_SHADE(EX, CN, MB) k_shade<EX, CN, MB><<<g, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_hit_colour, job.d_hit_lit, job.d_stats)
    if (job.precision == RT_FP64_EXACT) {
        if (count) LAUNCH_SHADE(true, true, 1); else LAUNCH_SHADE(true, false, 1);
    } else if (job.shade32) {          // opt-in fp32 shading (RT_B200_SHADE32=1): see shade_hit_fast32
        if (count) k_shade<false, true, 1, true><<<g, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_hit_colour, job.d_hit_lit, job.d_stats);
        else k_shade<false, false, 8, true><<<g, 128, 0, st>>>(fp, sc, job.fr, brute, job.d_hits, job.d_counters, job.d_hit_colour, job.d_hit_lit, job.d_stats);
    } else if (count) {
        LAUNCH_SHADE(false, true, 1);
    } else {
        if (job.shade_minb >= 8) LAUNCH_SHADE(false, false, 8);
        else if (job.shade_minb >= 6) LAUNCH_SHADE(false, false, 6);
        else if (job.shade_minb >= 5) LAUNCH_SHADE(false, false, 5);
        else LAUNCH_SHADE(false, false, 1);
    }
#undef LAUNCH_SHADE
    return cudaGetLastError();
}

cudaError_t rt_launch_fused(const RenderJob &job, cudaStream_t st) {
    const FrameParams &fp = job.fp;
    SceneView sc; sc.meshes = job.d_meshes; sc.objs = job.d_objs; sc.lights = job.d_lights; sc.trav_smem = nullptr; sc.tune = job.d_tune; sc.tune_host = job.d_tune_host; sc.defer = 0;
    const uint32_t nrays = ray_count(fp);
    if (nrays == 0) return cudaSuccess;
    const bool count = (job.flags & RT_FLAG_COUNT) != 0;
    const bool rowmajor = (job.flags & RT_LAYOUT_ROWMAJOR) != 0, sdl = (job.flags & RT_QUANT_RGBA16_SHR8) != 0;
    if (job.sm_affine && job.d_chunk_state) {
        // persistent SM-affine form: SMs x resident CTAs, work handed out in screen-local chunks (see k_frame_fused_sm)
        const ChunkGrid g = chunk_grid(fp);
        if ((size_t)2 + RT_MAX_SM + g.n_chunks > job.chunk_state_words || job.sm_count > RT_MAX_SM) return cudaErrorInvalidValue;
        cudaError_t e = cudaMemsetAsync(job.d_chunk_state, 0, sizeof(unsigned int) * ((size_t)2 + RT_MAX_SM + g.n_chunks), st);
        if (e != cudaSuccess) return e;
#define LAUNCH_FUSED_SM(CN, MB) do { prefer_l1(k_frame_fused_sm<CN, MB>);                                                             \
        k_frame_fused_sm<CN, MB><<<(unsigned)job.sm_count * (unsigned)(MB), 128, 0, st>>>(fp, sc, job.fr, rowmajor, sdl, job.d_counters, job.d_chunk_state, \
                                                    (unsigned)job.sm_count, job.d_out_rgb, job.d_out_colour, job.d_out_face, job.d_out_obj, job.d_out_lit, job.d_stats); } while (0)
        if (count) LAUNCH_FUSED_SM(true, 4);
        else if (job.shade_minb >= 9) LAUNCH_FUSED_SM(false, 9);
        else LAUNCH_FUSED_SM(false, 8);
#undef LAUNCH_FUSED_SM
        return cudaGetLastError();
    }
    const unsigned nblk = (nrays + RT_FUSED_THREADS - 1) / RT_FUSED_THREADS;
    // the traversal kernels use no shared memory: ask for the whole unified L1 / shared array as L1 (once per kernel and device)
    // the any-hit loop is a compile-time property of the variant: the plain loop when the (self-tuned or fixed) bound says 1
#define LAUNCH_FUSED_D(CN, MB, DF) do { prefer_l1(k_frame_fused<CN, MB, DF>);                                                          \
        k_frame_fused<CN, MB, DF><<<nblk, RT_FUSED_THREADS, 0, st>>>(fp, sc, job.fr, rowmajor, sdl, job.d_counters, job.d_out_rgb, job.d_out_colour, \
                                                    job.d_out_face, job.d_out_obj, job.d_out_lit, job.d_stats); } while (0)
#define LAUNCH_FUSED(CN, MB) do { if (job.any_plain) LAUNCH_FUSED_D(CN, MB, false); else LAUNCH_FUSED_D(CN, MB, true); } while (0)
    if (count) LAUNCH_FUSED(true, 1);
    else if (job.shade_minb >= 10) LAUNCH_FUSED(false, 10);
    else if (job.shade_minb >= 9) LAUNCH_FUSED(false, 9);
    else if (job.shade_minb >= 8) LAUNCH_FUSED(false, 8);
    else if (job.shade_minb >= 7) LAUNCH_FUSED(false, 7);
    else if (job.shade_minb >= 6) LAUNCH_FUSED(false, 6);
    else LAUNCH_FUSED(false, 1);
#undef LAUNCH_FUSED
#undef LAUNCH_FUSED_D
    return cudaGetLastError();
}

cudaError_t rt_launch_write(const RenderJob &job, cudaStream_t st) {
    if (ray_count(job.fp) == 0) return cudaSuccess;
    const bool rowmajor = (job.flags & RT_LAYOUT_ROWMAJOR) != 0, sdl = (job.flags & RT_QUANT_RGBA16_SHR8) != 0;
    k_write<<<(unsigned)job.sm_count * 8u, 256, 0, st>>>(job.fp, job.d_hits, job.d_counters, job.d_hit_colour, job.d_hit_lit,
                                                        rowmajor, sdl, job.d_out_rgb, job.d_out_colour, job.d_out_face,
                                                        job.d_out_obj, job.d_out_lit);
    return cudaGetLastError();
}

cudaError_t rt_launch_untile(const uint8_t *d_tiles, uint8_t *d_out, int W, int H, int tile_w, int tile_h, int n_ranks,
                             int tiles_per_rank_max, bool rowmajor, cudaStream_t st) {
    const int tiles_x = (W + tile_w - 1) / tile_w, tiles_y = (H + tile_h - 1) / tile_h;
    const size_t smem = (size_t)tile_w * tile_h * 3;
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    k_untile<<<(unsigned)(tiles_x * tiles_y), 128, smem, st>>>(d_tiles, d_out, W, H, tile_w, tile_h, n_ranks, tiles_per_rank_max, rowmajor);
    return cudaGetLastError();
}

cudaError_t rt_launch_clear_tiles(const FrameParams &fp, bool rowmajor, uint8_t *d_out, cudaStream_t st) {
    if (fp.n_local_tiles <= 0) return cudaSuccess;
    k_clear_tiles<<<(unsigned)fp.n_local_tiles, 128, 0, st>>>(fp, rowmajor, d_out);
    return cudaGetLastError();
}

cudaError_t rt_launch_push_tiles(const FrameParams &fp, bool rowmajor, const uint8_t *d_tiles, uint8_t *d_out, const PushSync &sy,
                                 int sm_count, cudaStream_t st) {
    if (fp.n_local_tiles <= 0) return cudaSuccess;
    const size_t smem = (size_t)fp.tile_w * fp.tile_h * 3;
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    const int per_sm = sy.debug & 8u ? 1024 : 2;
    const unsigned grid = (unsigned)(sy.slot ? std::min(fp.n_local_tiles, per_sm * sm_count) : fp.n_local_tiles);
    k_push_tiles<<<grid, 128, smem, st>>>(fp, rowmajor, d_tiles, d_out, sy);
    return cudaGetLastError();
}

cudaError_t rt_launch_frame_complete(uint32_t *d_slot, uint32_t seq, int n_ranks, bool release, uint32_t timeout_ms, cudaStream_t st) {
    k_frame_complete<<<1, 32, 0, st>>>(d_slot, seq, n_ranks, release ? 1 : 0, timeout_ms);
    return cudaGetLastError();
}

cudaError_t rt_launch_wait_consumed(uint32_t *d_slot, uint32_t seq, uint32_t rank, uint32_t timeout_ms, cudaStream_t st) {
    k_wait_consumed<<<1, 1, 0, st>>>(d_slot, seq, rank, timeout_ms);
    return cudaGetLastError();
}

cudaError_t rt_launch_signal_arrive(uint32_t *d_slot, uint32_t seq, uint32_t rank, cudaStream_t st) {
    k_signal_arrive<<<1, 1, 0, st>>>(d_slot, seq, rank);
    return cudaGetLastError();
}

cudaError_t rt_launch_frame_release(uint32_t *d_slot, uint32_t seq, cudaStream_t st) {
    k_frame_release<<<1, 1, 0, st>>>(d_slot, seq);
    return cudaGetLastError();
}
