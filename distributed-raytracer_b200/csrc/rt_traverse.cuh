// rt_traverse.cuh -- per-ray BVH traversal + scene-level trace/shade device functions.
//
// Everything here is RT_HD (host+device) and free of warp intrinsics so that tests/emul can run
// the identical code on the CPU; the persistent warp-cooperative primary kernel in render.cu
// re-uses the node/triangle primitives below inside its own restructured loop.
//
// Reference semantics implemented (paths under /root/reference/):
//   shared/state/object.go:63-110   per-object nearest hit, metric Len(o_local - ip), strict '<'
//   worker/shared/tracer/tracer.go:27-50  world nearest, metric Len(ip_world - Cam.Pos), strict '<'
//   worker/shared/tracer/tracer.go:53-77  phong with full nearest-hit shadow queries
// The R-tree candidate search (object.go:76, box.go:29-69) is replaced by a BVH that may only CULL:
// in FP64_EXACT mode boxes are padded and the cull bound is inflated so that every face the
// reference's Moller-Trumbore test accepts and that could win is still tested, and the winner is
// the lexicographic (distance, face index) minimum == "first in enumeration order, strict <".
#pragma once
#include "rt_exact.cuh"

struct alignas(16) rt_d2 { double x, y; };

struct SceneView {
    const MeshView *meshes;
    const ObjRec   *objs;
    const LightRec *lights;
    // Shared-memory column of this thread for the bottom of the traversal stack (device kernels; nullptr = local memory only).
    // Entry k of the thread lives at trav_smem[k * RT_SMEM_STACK_STRIDE]: lanes with different stack depths still hit 32 different
    // banks, so a push / pop is ONE shared-memory wavefront per warp instead of up to 32 local-memory wavefronts through L1
    // (ncu, fused kernel: local loads + stores were 28 % of the LSU wavefronts and 40 % of the L1 sectors).
    uint32_t *trav_smem;
    // Scene-level tuning words in device memory (null in the CPU replay): [0] shadow rays counted, [1] of which occluded,
    // [2] leaves an any-hit ray postpones before it tests them.  The render kernels keep [0] / [1] (one CTA in four samples its
    // rays) and re-derive [2] every 16 384 sampled rays (render.cu: tune_any_defer): the answer of an any-hit ray does not depend on the order of its tests, so the knob can
    // move under running frames without changing a bit of the output.
    unsigned int *tune;
    unsigned int *tune_host;   // host-mapped copy of word [2]: lets the HOST pick the kernel variant without a synchronisation
    uint32_t defer;            // the bound the current warp works with (read from tune[2] once per pixel block; 0 = the mesh's own)
};

// MEASURED SLOWER (profiles/r02_kernel_ab.txt: fused frame 0.627 -> 0.726 ms pipelined, K2 +15 %, K4 +9 %): the depth test and the
// second address stream cost more issue slots than the saved L1 wavefronts give back, and 72 KB of shared memory per SM come out
// of the L1.  Kept as a compile-time experiment (-DRT_SMEM_STACK=1), off by default.
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 0
#endif
#define RT_SMEM_STACK_STRIDE 128          // threads per CTA of every traversal kernel
#define RT_SMEM_STACK_WORDS 16            // 4-byte words per thread: 16 refs (any-hit) or 8 x (ref, t) (nearest-hit)

#ifndef RT_COUNT   // (tests/emul redefines it to log the node / leaf visit sequence of a ray: tools/simulate_warp_divergence_cpu.py)
#define RT_COUNT(c, field, n) do { if (c) (c)->field += (n); } while (0)
#endif

// Measured on B200 (profiles/r01_ab_traversal_shape_mt_sync_minmax.txt), config 3:
//  * loop shape: the plain "if internal node ... else leaf" loop below beats a while-while loop with postponed
//    leaves (no gain) and a while-while loop with a warp vote (30 % slower) -- kept simple.
//  * the light loop of shade_hit needs an explicit __syncwarp() per light: without it the lanes of a warp never
//    reconverge after the first shadow ray and the rest of the kernel runs at ~4 active lanes.
#if defined(__CUDA_ARCH__)
#define RT_SYNCWARP() __syncwarp()
#elif !defined(RT_SYNCWARP)
#define RT_SYNCWARP() do { } while (0)
#endif

RT_HD rt_f4 ld_f4(const rt_f4 *p) {
#if defined(__CUDA_ARCH__)
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    rt_f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}
// one 32-byte vector through the read-only path: LDG.E.256 on sm_100a (needs 32-byte alignment)
RT_HD rt_f8 ld_f8(const rt_f4 *p) {
#if defined(__CUDA_ARCH__) && RT_LD256
    rt_f8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
#else
    rt_f8 r; r.a = ld_f4(p); r.b = ld_f4(p + 1); return r;
#endif
}
// Cache policy (profiles/r02_kernel_ab.txt): RT_TRI_NOALLOC (default) loads the triangle records of a leaf with L1::no_allocate -- they are
// used once per ray, and the lines they would evict hold BVH nodes: -0.7 % on the pipelined C3 frame, small but repeatable to 0.01 %.
// RT_NODE_EVICT_LAST (marking the node lines evict_last in L1) changes nothing and stays off.
#ifndef RT_TRI_NOALLOC
#define RT_TRI_NOALLOC 1
#endif
#ifndef RT_NODE_EVICT_LAST
#define RT_NODE_EVICT_LAST 0
#endif
RT_HD rt_f8 ld_f8_tri(const rt_f4 *p) {
#if defined(__CUDA_ARCH__) && RT_LD256 && RT_TRI_NOALLOC
    rt_f8 r;
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
#else
    return ld_f8(p);
#endif
}
RT_HD rt_f8 ld_f8_node(const rt_f4 *p) {
#if defined(__CUDA_ARCH__) && RT_LD256 && RT_NODE_EVICT_LAST
    rt_f8 r;
    asm("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
#else
    return ld_f8(p);
#endif
}
RT_HD rt_i4 ld_i4(const rt_i4 *p) {
#if defined(__CUDA_ARCH__)
    int4 v = __ldg(reinterpret_cast<const int4 *>(p));
    rt_i4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}
RT_HD rt_d2 ld_d2(const double *p) {
#if defined(__CUDA_ARCH__)
    double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    rt_d2 r; r.x = v.x; r.y = v.y; return r;
#else
    rt_d2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}

// one node visit = 7 * RT_W/4 vector loads (the pad array is never fetched)
struct NodeRegs { float lox[RT_W], hix[RT_W], loy[RT_W], hiy[RT_W], loz[RT_W], hiz[RT_W]; int ref[RT_W]; };
RT_HD void ld_node(const RtNode *n, NodeRegs &r) {
    const rt_f4 *q = reinterpret_cast<const rt_f4 *>(n);
    float *dst = r.lox;                       // the six box arrays are contiguous in both structs
#pragma unroll
    for (int k = 0; k < 6 * RT_W / 4; k += 2) {
        rt_f8 v = ld_f8(q + k);
        dst[4 * k] = v.a.x; dst[4 * k + 1] = v.a.y; dst[4 * k + 2] = v.a.z; dst[4 * k + 3] = v.a.w;
        dst[4 * k + 4] = v.b.x; dst[4 * k + 5] = v.b.y; dst[4 * k + 6] = v.b.z; dst[4 * k + 7] = v.b.w;
    }
    const rt_i4 *qi = reinterpret_cast<const rt_i4 *>(n->ref);
#pragma unroll
    for (int k = 0; k < RT_W / 4; k++) {
        rt_i4 v = ld_i4(qi + k);
        r.ref[4 * k] = v.x; r.ref[4 * k + 1] = v.y; r.ref[4 * k + 2] = v.z; r.ref[4 * k + 3] = v.w;
    }
}

// ------------------------------------------------------------------------------------------------
// FP64_EXACT
// ------------------------------------------------------------------------------------------------
struct BestExact { bool has; double dist; TriHit h; int face; };

// object.go:88-105 for one candidate face
RT_HD void consider_exact(dv P1, dv P2, dv P3, int face, dv o, dv d, BestExact &b) {
    TriHit h;
    if (tri_intersection_exact(P1, P2, P3, o, d, h)) {                      // object.go:88
        double dist = dv_len(dv_sub(o, h.ip));                              // object.go:97
        if (!b.has || dist < b.dist || (dist == b.dist && face < b.face)) { // object.go:98 (+ lowest index on ties)
            b.has = true; b.dist = dist; b.h = h; b.face = face;
        }
    }
}

// conservative fp64 slab test of a padded fp32 box
RT_HD bool slab64(float lox, float hix, float loy, float hiy, float loz, float hiz, dv o, dv idir, double cull,
                  double &tin) {
    double t1 = ((double)lox - o.x) * idir.x, t2 = ((double)hix - o.x) * idir.x;
    double tmin = fmin(t1, t2), tmax = fmax(t1, t2);
    t1 = ((double)loy - o.y) * idir.y; t2 = ((double)hiy - o.y) * idir.y;
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    t1 = ((double)loz - o.z) * idir.z; t2 = ((double)hiz - o.z) * idir.z;
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    tin = tmin;
    return tmin <= tmax && tmax >= 0.0 && tmin <= cull;
}

RT_HD double cull_bound(const BestExact &b) { return b.has ? b.dist * (1.0 + 1e-9) + 1e-300 : HUGE_VAL; }

// Nearest hit of one mesh in object space.  brute => every face in index order (no BVH).
RT_HD void mesh_nearest_exact(const MeshView &m, dv o, dv d, BestExact &best, bool brute, Counters *cnt) {
    if (m.n_faces == 0) return;
    if (brute) {
        for (uint32_t f = 0; f < m.n_faces; f++) {
            const double *p = m.fverts + 9 * (size_t)f;
            consider_exact(dv_ld(p), dv_ld(p + 3), dv_ld(p + 6), (int)f, o, d, best);
        }
        RT_COUNT(cnt, tris, m.n_faces);
        return;
    }
    dv idir = dv_make(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
    double tin;
    if (!slab64(m.lo[0], m.hi[0], m.lo[1], m.hi[1], m.lo[2], m.hi[2], o, idir, cull_bound(best), tin)) return;
    int stack[RT_STACK_SIZE];
    double stack_t[RT_STACK_SIZE];
    int sp = 0;
    int ref = m.root_ref;
    for (;;) {
        if (ref >= 0) {
            NodeRegs n;
            ld_node(m.nodes + ref, n);
            RT_COUNT(cnt, nodes, 1);
            const double cull = cull_bound(best);
            int next = RT_EMPTY;
            double tnext = 0.0;
#pragma unroll
            for (int k = 0; k < RT_W; k++) {
                double t;
                bool h = slab64(n.lox[k], n.hix[k], n.loy[k], n.hiy[k], n.loz[k], n.hiz[k], o, idir, cull, t) && n.ref[k] != RT_EMPTY;
                if (h) {
                    // continue with the nearest child, push the others (with their entry distance)
                    if (next == RT_EMPTY) { next = n.ref[k]; tnext = t; }
                    else if (t < tnext) { stack[sp] = next; stack_t[sp] = tnext; sp++; next = n.ref[k]; tnext = t; }
                    else { stack[sp] = n.ref[k]; stack_t[sp] = t; sp++; }
                }
            }
            if (next != RT_EMPTY) { ref = next; continue; }
        } else {
            uint32_t first = rt_leaf_first(ref), count = rt_leaf_count(ref);
            RT_COUNT(cnt, tris, count);
            for (uint32_t k = first; k < first + count; k++) {
                const double *p = m.tri64 + 10 * (size_t)k;
                rt_d2 q0 = ld_d2(p), q1 = ld_d2(p + 2), q2 = ld_d2(p + 4), q3 = ld_d2(p + 6), q4 = ld_d2(p + 8);
                consider_exact(dv_make(q0.x, q0.y, q1.x), dv_make(q1.y, q2.x, q2.y), dv_make(q3.x, q3.y, q4.x),
                               (int)q4.y, o, d, best);
            }
        }
        // pop; entries that a closer hit has made irrelevant since they were pushed are skipped (conservative bound)
        do {
            if (sp == 0) return;
            --sp;
            ref = stack[sp];
        } while (stack_t[sp] > cull_bound(best));
    }
}

struct TraceExact { dv ip; int face, obj; double r1, r2, r3; };

// tracer.go:27-50 (objects visited in scene order; ties of the camera distance keep the first)
RT_HD bool trace_exact(const SceneView &sc, const FrameParams &fp, dv o, dv d, bool brute, TraceExact &out,
                       Counters *cnt) {
    bool exists = false;
    double nd = 0.0;
    dv cam = dv_ld(fp.cam_pos);
    for (int k = 0; k < fp.n_obj; k++) {
        const ObjRec &ob = sc.objs[k];
        if (!(ob.pos[0] == ob.pos[0])) continue;          // NaN position: the object is absent from this order (rt_b200.h)
        dv pos = dv_ld(ob.pos);
        const MeshView &mv = sc.meshes[ob.mesh];
        if (!brute) {                                                                    // tracer.go:32: box.go's predicate on the object's bounds
            dv bmin, bmax;
            object_box_exact(pos, dv_ld(mv.vlo), dv_ld(mv.vhi), bmin, bmax);
            if (!box_intersect_exact(bmin, bmax, o, d)) continue;
        }
        BestExact b; b.has = false; b.dist = 0.0; b.face = -1;
        mesh_nearest_exact(mv, dv_sub(o, pos), d, b, brute, cnt);                        // object.go:71
        if (b.has) {
            dv ipw = dv_add(b.h.ip, pos);                                                // object.go:109
            double dist = dv_len(dv_sub(ipw, cam));                                      // tracer.go:38
            if (!exists || dist < nd) {                                                  // tracer.go:39
                exists = true; nd = dist;
                out.ip = ipw; out.face = b.face; out.obj = k; out.r1 = b.h.r1; out.r2 = b.h.r2; out.r3 = b.h.r3;
            }
        }
    }
    return exists;
}

// ------------------------------------------------------------------------------------------------
// FP32_FAST
// ------------------------------------------------------------------------------------------------
struct RayF { float ox, oy, oz, dx, dy, dz, ix, iy, iz, ax, ay, az; };   // i = 1/d, a = o/d

RT_HD uint32_t rt_float_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RT_HD float rt_bits_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

RT_HD float safe_rcp(float d) {
    const float eps = 8.2718061e-25f;   // 2^-80
    float ad = fabsf(d) > eps ? d : copysignf(eps, d);
    return 1.0f / ad;
}
RT_HD RayF make_rayf(float ox, float oy, float oz, float dx, float dy, float dz) {
    RayF r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = safe_rcp(dx); r.iy = safe_rcp(dy); r.iz = safe_rcp(dz);
    r.ax = ox * r.ix; r.ay = oy * r.iy; r.az = oz * r.iz;
    return r;
}

// slab test of one child box; returns entry distance in tin
RT_HD bool slab32(float lox, float hix, float loy, float hiy, float loz, float hiz, const RayF &r, float tbest,
                  float &tin) {
    float x0 = fmaf(lox, r.ix, -r.ax), x1 = fmaf(hix, r.ix, -r.ax);
    float y0 = fmaf(loy, r.iy, -r.ay), y1 = fmaf(hiy, r.iy, -r.ay);
    float z0 = fmaf(loz, r.iz, -r.az), z1 = fmaf(hiz, r.iz, -r.az);
    float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tbest));
    tin = tmin;
    return tmin <= tmax;
}

// fp32 Moller-Trumbore, same acceptance rules as triangle.go:37-77 (two-sided, inclusive, t >= 0).  The range tests
// are done on the un-normalised numerators against |det| so that the reciprocal is only paid by accepted hits.
RT_HD bool tri_test32(rt_f4 v0, rt_f4 v1, rt_f4 v2, const RayF &r, float &t, float &u, float &v) {
    float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
    float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
    float px = fmaf(r.dy, e2z, -(r.dz * e2y)), py = fmaf(r.dz, e2x, -(r.dx * e2z)), pz = fmaf(r.dx, e2y, -(r.dy * e2x));
    float det = fmaf(e1x, px, fmaf(e1y, py, e1z * pz));
    float ad = fabsf(det);
    float tx = r.ox - v0.x, ty = r.oy - v0.y, tz = r.oz - v0.z;
    float un = copysignf(1.0f, det) * fmaf(tx, px, fmaf(ty, py, tz * pz));
    if (!(un >= 0.0f && un <= ad) || det == 0.0f) return false;
    float qx = fmaf(ty, e1z, -(tz * e1y)), qy = fmaf(tz, e1x, -(tx * e1z)), qz = fmaf(tx, e1y, -(ty * e1x));
    float vn = copysignf(1.0f, det) * fmaf(r.dx, qx, fmaf(r.dy, qy, r.dz * qz));
    if (!(vn >= 0.0f && un + vn <= ad)) return false;
    float tn = copysignf(1.0f, det) * fmaf(e2x, qx, fmaf(e2y, qy, e2z * qz));
    if (!(tn >= 0.0f)) return false;
    float inv = 1.0f / ad;
    t = tn * inv; u = un * inv; v = vn * inv;
    return true;
}

// The 4 slab tests of one node visit.  RT_OCTANT (default): the sign of each direction component says which of the two planes
// of an axis the ray enters through, so the near / far plane arrays are picked by ADDRESS (the node keeps lo and hi of an axis in
// adjacent 16-byte vectors) and the per-child min/max pair of the classic slab test disappears: 6 FFMA + 2 three-input min/max
// per child instead of 6 FFMA + 6 min/max + 2.  Bit-identical to slab32: lo <= hi and x -> fma(x, i, -a) is monotone, so
// min(x0, x1) is x0 for i >= 0 and x1 for i < 0.
#ifndef RT_OCTANT
#define RT_OCTANT 1
#endif
struct NodeHits { float t[RT_W]; int ref[RT_W]; bool hit[RT_W]; };
RT_HD void node_test32(const RtNode *nd, const RayF &r, float tbest, NodeHits &h) {
#if RT_LD256 && RT_W == 4
    // lo[4] and hi[4] of an axis are one 32-byte vector: 3 x LDG.256 + the refs = 4 load instructions per visit
    const rt_f4 *q = reinterpret_cast<const rt_f4 *>(nd);
    const rt_f8 bx = ld_f8_node(q), by = ld_f8_node(q + 2), bz = ld_f8_node(q + 4);
    const rt_i4 rf = ld_i4(reinterpret_cast<const rt_i4 *>(q + 6));
#define RT_CHILD(k, c)                                                                                                     \
    {                                                                                                                      \
        const float x0 = fmaf(bx.a.c, r.ix, -r.ax), x1 = fmaf(bx.b.c, r.ix, -r.ax);                                        \
        const float y0 = fmaf(by.a.c, r.iy, -r.ay), y1 = fmaf(by.b.c, r.iy, -r.ay);                                        \
        const float z0 = fmaf(bz.a.c, r.iz, -r.az), z1 = fmaf(bz.b.c, r.iz, -r.az);                                        \
        const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));                           \
        const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tbest));                          \
        h.t[k] = tn; h.ref[k] = rf.c; h.hit[k] = tn <= tf && rf.c != RT_EMPTY;                                             \
    }
    RT_CHILD(0, x) RT_CHILD(1, y) RT_CHILD(2, z) RT_CHILD(3, w)
#undef RT_CHILD
#elif RT_OCTANT && RT_W == 4
    const rt_f4 *q = reinterpret_cast<const rt_f4 *>(nd);
    const int sx = r.ix < 0.0f, sy = r.iy < 0.0f, sz = r.iz < 0.0f;
    const rt_f4 nx = ld_f4(q + sx), fx = ld_f4(q + (sx ^ 1));
    const rt_f4 ny = ld_f4(q + 2 + sy), fy = ld_f4(q + 2 + (sy ^ 1));
    const rt_f4 nz = ld_f4(q + 4 + sz), fz = ld_f4(q + 4 + (sz ^ 1));
    const rt_i4 rf = ld_i4(reinterpret_cast<const rt_i4 *>(q + 6));
#define RT_CHILD(k, c)                                                                                                     \
    {                                                                                                                      \
        const float tn = fmaxf(fmaxf(fmaf(nx.c, r.ix, -r.ax), fmaf(ny.c, r.iy, -r.ay)), fmaxf(fmaf(nz.c, r.iz, -r.az), 0.0f)); \
        const float tf = fminf(fminf(fmaf(fx.c, r.ix, -r.ax), fmaf(fy.c, r.iy, -r.ay)), fminf(fmaf(fz.c, r.iz, -r.az), tbest)); \
        h.t[k] = tn; h.ref[k] = rf.c; h.hit[k] = tn <= tf && rf.c != RT_EMPTY;                                             \
    }
    RT_CHILD(0, x) RT_CHILD(1, y) RT_CHILD(2, z) RT_CHILD(3, w)
#undef RT_CHILD
#else
    NodeRegs n;
    ld_node(nd, n);
#pragma unroll
    for (int k = 0; k < RT_W; k++) {
        h.ref[k] = n.ref[k];
        h.hit[k] = slab32(n.lox[k], n.hix[k], n.loy[k], n.hiy[k], n.loz[k], n.hiz[k], r, tbest, h.t[k]) && n.ref[k] != RT_EMPTY;
    }
#endif
}

// fp32 test of the triangles of one leaf of an any-hit ray: true at the first triangle with t <= tmax.
// RT_LEAF_PAIR (compile-time experiment, off): the records of the first TWO triangles are requested together (any-hit leaves hold <= 2
// by default), so a two-triangle leaf costs one load latency instead of two.  Measured (profiles/r02_kernel_ab.txt): K4 of the
// three-kernel pipeline 0.559 -> 0.550 ms, but the fused frame kernel -- which sits at its register limit (56) -- 0.605 -> 0.633 ms.
#ifndef RT_LEAF_PAIR
#define RT_LEAF_PAIR 0
#endif
RT_HD bool leaf_any_hit32(const MeshView &m, int ref, const RayF &r, float tmax, int &face, float &thit, Counters *cnt) {
    const uint32_t first = rt_leaf_first(ref), count = rt_leaf_count(ref);
    RT_COUNT(cnt, tris, count);
    uint32_t k = first;
#if RT_LEAF_PAIR && RT_TRI_LD256
    {
        const rt_f4 *p = m.tri32 + RT_TRI32_F4 * (size_t)first;
        const bool two = count >= 2u;
        const rt_f8 a01 = ld_f8_tri(p), a2x = ld_f8_tri(p + 2);
        rt_f8 b01 = a01, b2x = a2x;
        if (two) { b01 = ld_f8_tri(p + 4); b2x = ld_f8_tri(p + 6); }
        float t, u, v;
        if (tri_test32(a01.a, a01.b, a2x.a, r, t, u, v) && t <= tmax) { face = (int)rt_float_bits(a01.a.w); thit = t; return true; }
        if (two && tri_test32(b01.a, b01.b, b2x.a, r, t, u, v) && t <= tmax) { face = (int)rt_float_bits(b01.a.w); thit = t; return true; }
        k = first + 2u;
    }
#endif
    for (; k < first + count; k++) {
        const rt_f4 *p = m.tri32 + RT_TRI32_F4 * (size_t)k;
#if RT_TRI_LD256
        const rt_f8 t01 = ld_f8_tri(p), t2x = ld_f8_tri(p + 2);
        const rt_f4 v0 = t01.a, v1 = t01.b, v2 = t2x.a;
#else
        const rt_f4 v0 = ld_f4(p), v1 = ld_f4(p + 1), v2 = ld_f4(p + 2);
#endif
        float t, u, v;
        if (tri_test32(v0, v1, v2, r, t, u, v) && t <= tmax) {
            face = (int)rt_float_bits(v0.w);
            thit = t;
            return true;
        }
    }
    return false;
}

// Any-hit traversal with EVERY LEAF POSTPONED (RT_ANY_DEFER, the default for shadow rays).
// Why: 93 % of the shadow rays of config 3 are lit, i.e. the ray must visit every node and leaf its segment meets and the
// order is irrelevant.  In the plain "if node ... else leaf" loop the lanes of a warp sit at nodes and leaves at different
// times, so a warp issued 25 node bodies (22 active lanes) PLUS 15.5 leaf bodies with 6 active lanes per shadow ray (ncu source
// page, profiles/r02_ncu_c3_summary.txt) -- each a dependent load + arithmetic chain the whole warp waits for.  Here a ray first
// walks the internal nodes only (leaf children go to the far end of the same local array), then tests its postponed leaves
// back to back: the warp runs max-over-lanes(leaves per ray) leaf bodies instead of one per lane-and-time.  The answer
// (is anything closer than the light?) does not depend on the order, so frames are bit-identical to the plain loop.
// An occluded ray finds its occluder later than before, so the postponement is BOUNDED: once `defer` leaves are waiting they are
// tested before the walk goes on.  Measured (profiles/r02_kernel_ab.txt, ms per pipelined frame; bound 1 / 2 / 4 / 8 / 16 / none):
//   C3 (7 % of the shadow rays occluded):   0.708 / 0.664 / 0.627 / 0.612 / 0.610 / 0.611     (plain loop: 0.626)
//   C4 (10 M-triangle soup, mostly occluded): 27.1 / 25.2 / 24.2 / 24.9 / 27.6 / 33.2          (plain loop: 22.8)
// No single bound suits both, so the kernels measure the occluded share of their own shadow rays and pick 16, 8 or 1 for the rays that follow (SceneView::tune; 1 = the plain loop: the host launches the kernel variant compiled
// with it); RT_B200_ANY_DEFER fixes it instead.
#ifndef RT_ANY_DEFER
#define RT_ANY_DEFER 1
#endif
// experiment: a lit ray visits every node it pushes, so the line of a pushed child can be requested at push time
#ifndef RT_PUSH_PREFETCH
#define RT_PUSH_PREFETCH 0
#endif
#if defined(__CUDA_ARCH__) && RT_PUSH_PREFETCH == 1
#define RT_PREFETCH_NODE(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#elif defined(__CUDA_ARCH__) && RT_PUSH_PREFETCH == 2
#define RT_PREFETCH_NODE(p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p))
#else
#define RT_PREFETCH_NODE(p) do { } while (0)
#endif
RT_HD bool mesh_any_hit_deferred(const MeshView &m, const RayF &r, float &tmax, int &face, Counters *cnt, uint32_t defer = 0) {
    int stack[RT_STACK_SIZE];            // internal nodes: [0, sp) growing up; postponed leaves: [lp, RT_STACK_SIZE) growing down
    int sp = 0, lp = RT_STACK_SIZE;
    const uint32_t bound = defer ? defer : m.any_defer;
    const int lp_drain = RT_STACK_SIZE - (int)(bound < 4u ? 4u : bound);      // a kernel variant without the plain loop treats 1 .. 3 as 4
    int ref = m.root_ref_any;
    bool more = true;
    if (ref < 0) { stack[--lp] = ref; more = false; }
#if RT_USE_NODEQ
    const RayQ rq = make_rayq(m, r);
#endif
    do {
        while (more) {
            NodeHits n;
#if RT_USE_NODEQ
            node_test32q(m.qnodes_any + ref, rq, tmax, n);
#else
            node_test32(m.nodes_any + ref, r, tmax, n);
#endif
            RT_COUNT(cnt, nodes, 1);
            int next = RT_EMPTY;
#pragma unroll
            for (int k = 0; k < RT_W; k++) {
                if (n.hit[k]) {
                    const int c = n.ref[k];
                    if (c < 0) stack[--lp] = c;
                    else if (next == RT_EMPTY) next = c;
                    else { stack[sp++] = c; RT_PREFETCH_NODE(m.nodes_any + c); }
                }
            }
            if (next != RT_EMPTY) ref = next;
            else if (sp > 0) ref = stack[--sp];
            else more = false;
            // enough leaves are waiting (or the array is nearly full): test them before walking on, so that an OCCLUDED ray meets its
            // occluder after at most any_defer + RT_W - 1 postponed leaves instead of after its whole node walk
            if (lp <= lp_drain || lp - sp < 2 * RT_W) break;
        }
        while (lp < RT_STACK_SIZE) {
            float thit;
            if (leaf_any_hit32(m, stack[lp++], r, tmax, face, thit, cnt)) { tmax = thit; return true; }
        }
    } while (more);
    return false;
}

// ---- quantised 64-byte nodes (rt_types.h: RtNodeQ) ----
struct alignas(32) rt_u8 { uint32_t w[8]; };
RT_HD rt_u8 ld_u8(const void *p) {
#if defined(__CUDA_ARCH__)
    rt_u8 r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
#else
    rt_u8 r; memcpy(&r, p, 32); return r;
#endif
}
// 2^23 + (low or high 16-bit field of w) as a float: the field becomes the low mantissa bits of 0x4B000000 (one PRMT)
RT_HD float q16f(uint32_t w, bool high) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, high ? 0x7632u : 0x7610u));
#else
    return rt_bits_float(0x4B000000u | (high ? (w >> 16) : (w & 0xffffu)));
#endif
}
// the ray in the grid of a mesh's quantised nodes: plane q of an axis is met at t = (2^23 + q) * s + c
struct RayQ { float sx, sy, sz, cx, cy, cz; };
RT_HD RayQ make_rayq(const MeshView &m, const RayF &r) {
    RayQ q;
    q.sx = m.qstep[0] * r.ix; q.sy = m.qstep[1] * r.iy; q.sz = m.qstep[2] * r.iz;
    q.cx = fmaf(-8388608.0f, q.sx, (m.qorg[0] - r.ox) * r.ix);
    q.cy = fmaf(-8388608.0f, q.sy, (m.qorg[1] - r.oy) * r.iy);
    q.cz = fmaf(-8388608.0f, q.sz, (m.qorg[2] - r.oz) * r.iz);
    return q;
}
RT_HD void node_test32q(const RtNodeQ *nd, const RayQ &q, float tbest, NodeHits &h) {
    const rt_u8 a = ld_u8(nd), b = ld_u8(reinterpret_cast<const char *>(nd) + 32);     // a: x, y planes; b: z planes, refs
#define RT_CHILDQ(k)                                                                                                       \
    {                                                                                                                      \
        const float x0 = fmaf(q16f(a.w[(k) >> 1], (k) & 1), q.sx, q.cx), x1 = fmaf(q16f(a.w[2 + ((k) >> 1)], (k) & 1), q.sx, q.cx);   \
        const float y0 = fmaf(q16f(a.w[4 + ((k) >> 1)], (k) & 1), q.sy, q.cy), y1 = fmaf(q16f(a.w[6 + ((k) >> 1)], (k) & 1), q.sy, q.cy); \
        const float z0 = fmaf(q16f(b.w[(k) >> 1], (k) & 1), q.sz, q.cz), z1 = fmaf(q16f(b.w[2 + ((k) >> 1)], (k) & 1), q.sz, q.cz);   \
        const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));                           \
        const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tbest));                          \
        const int rf = (int)b.w[4 + (k)];                                                                                  \
        h.t[k] = tn; h.ref[k] = rf; h.hit[k] = tn <= tf && rf != RT_EMPTY;                                                 \
    }
    RT_CHILDQ(0) RT_CHILDQ(1) RT_CHILDQ(2) RT_CHILDQ(3)
#undef RT_CHILDQ
}
#if RT_NODEQ && RT_W == 4
#define RT_USE_NODEQ 1
#else
#define RT_USE_NODEQ 0
#endif

// Simple per-ray traversal of one mesh (object space).  ANY: stop at the first hit with t <= tbest.
// Nearest: winner = min t, ties -> lowest face index.  Returns true if something was accepted.
// ANY rays walk the mesh's any-hit hierarchy (nodes_any: same triangles, its own leaf size).
template <bool ANY, bool DEFER = (RT_ANY_DEFER != 0)>
RT_HD bool mesh_trace_fast(const MeshView &m, const RayF &r, float &tbest, int &face, Counters *cnt, uint32_t *smem = nullptr, uint32_t defer = 0) {
    if (m.n_faces == 0) return false;
    float tin;
    if (!slab32(m.lo[0], m.hi[0], m.lo[1], m.hi[1], m.lo[2], m.hi[2], r, tbest, tin)) return false;
    // DEFER = false is the plain loop below: the right one for scenes whose shadow rays are mostly occluded (a bound of 1).  It is a
    // COMPILE-TIME choice per kernel variant, picked by the host from the self-tuned bound: both loops in one kernel cost the fused
    // frame of C3 2.7 % (registers), profiles/r02_kernel_ab.txt
    if (ANY && DEFER) return mesh_any_hit_deferred(m, r, tbest, face, cnt, defer);
    bool found = false;
    int stack[RT_STACK_SIZE];
    float stack_t[ANY ? 1 : RT_STACK_SIZE];
    int sp = 0;
    // the first SK entries live in shared memory (see SceneView::trav_smem), deeper ones in the local arrays above
#if RT_SMEM_STACK
    const int SK = smem ? (ANY ? RT_SMEM_STACK_WORDS : RT_SMEM_STACK_WORDS / 2) : 0;
#else
    const int SK = 0;
    (void)smem;
#endif
#define RT_PUSH(REF, T)                                                                                                    \
    do {                                                                                                                   \
        if (sp < SK) {                                                                                                     \
            smem[sp * RT_SMEM_STACK_STRIDE] = (uint32_t)(REF);                                                             \
            if (!ANY) smem[(SK + sp) * RT_SMEM_STACK_STRIDE] = rt_float_bits(T);                                           \
        } else {                                                                                                           \
            stack[sp] = (REF);                                                                                             \
            if (!ANY) stack_t[sp] = (T);                                                                                   \
        }                                                                                                                  \
        sp++;                                                                                                              \
    } while (0)
#if RT_USE_NODEQ
    const RtNodeQ *nodes = ANY ? m.qnodes_any : m.qnodes;
    const RayQ rq = make_rayq(m, r);
#else
    const RtNode *nodes = ANY ? m.nodes_any : m.nodes;
#endif
    int ref = ANY ? m.root_ref_any : m.root_ref;
    for (;;) {
        if (ref >= 0) {
            NodeHits n;
#if RT_USE_NODEQ
            node_test32q(nodes + ref, rq, tbest, n);
#else
            node_test32(nodes + ref, r, tbest, n);
#endif
            RT_COUNT(cnt, nodes, 1);
            int next = RT_EMPTY;
            float tnext = 0.0f;
#pragma unroll
            for (int k = 0; k < RT_W; k++) {
                if (n.hit[k]) {
                    const float t = n.t[k];
                    if (next == RT_EMPTY) { next = n.ref[k]; tnext = t; }
                    else if (!ANY && t < tnext) { RT_PUSH(next, tnext); next = n.ref[k]; tnext = t; }
                    else RT_PUSH(n.ref[k], t);
                }
            }
            if (next != RT_EMPTY) { ref = next; continue; }
        } else {
            uint32_t first = rt_leaf_first(ref), count = rt_leaf_count(ref);
            RT_COUNT(cnt, tris, count);
            for (uint32_t k = first; k < first + count; k++) {
                const rt_f4 *p = m.tri32 + RT_TRI32_F4 * (size_t)k;
#if RT_TRI_LD256
                const rt_f8 t01 = ld_f8_tri(p), t2x = ld_f8_tri(p + 2);
                const rt_f4 v0 = t01.a, v1 = t01.b, v2 = t2x.a;
#else
                rt_f4 v0 = ld_f4(p), v1 = ld_f4(p + 1), v2 = ld_f4(p + 2);
#endif
                float t, u, v;
                if (tri_test32(v0, v1, v2, r, t, u, v)) {
#if defined(__CUDA_ARCH__)
                    int f = __float_as_int(v0.w);
#else
                    int f; { float w = v0.w; memcpy(&f, &w, 4); }
#endif
                    if (ANY) { if (t <= tbest) { face = f; tbest = t; return true; } }
                    else if (t < tbest || (t == tbest && (!found || f < face))) { tbest = t; face = f; found = true; }
                }
            }
        }
        // pop (nearest-hit mode skips entries that start beyond the current best hit)
        float tpop = 0.0f;
        do {
            if (sp == 0) return found;
            --sp;
            if (sp < SK) {
                ref = (int)smem[sp * RT_SMEM_STACK_STRIDE];
                if (!ANY) tpop = rt_bits_float(smem[(SK + sp) * RT_SMEM_STACK_STRIDE]);
            } else {
                ref = stack[sp];
                if (!ANY) tpop = stack_t[sp];
            }
        } while (!ANY && tpop > tbest);
    }
#undef RT_PUSH
}

struct TraceFast { int face, obj; float t; };

// World-level nearest hit in fp32.  from_camera: primary ray (metric == t); otherwise (shadow ray in a
// multi-object scene) objects are ranked by the hit's distance to the camera, tracer.go:38.
RT_HD bool trace_fast_nearest(const SceneView &sc, const FrameParams &fp, dv o, dv d, bool from_camera, TraceFast &out,
                              Counters *cnt) {
    bool exists = false;
    float best_metric = 0.0f;
    for (int k = 0; k < fp.n_obj; k++) {
        const ObjRec &ob = sc.objs[k];
        if (!(ob.pos[0] == ob.pos[0])) continue;          // NaN position: the object is absent from this order (rt_b200.h)
        RayF r = make_rayf((float)(o.x - ob.pos[0]), (float)(o.y - ob.pos[1]), (float)(o.z - ob.pos[2]), (float)d.x,
                           (float)d.y, (float)d.z);
        float t = 3.0e38f;
        int face = -1;
        if (mesh_trace_fast<false>(sc.meshes[ob.mesh], r, t, face, cnt, sc.trav_smem)) {
            float metric = t;
            if (!from_camera) {
                float px = (float)(o.x + d.x * t - fp.cam_pos[0]), py = (float)(o.y + d.y * t - fp.cam_pos[1]),
                      pz = (float)(o.z + d.z * t - fp.cam_pos[2]);
                metric = sqrtf(px * px + py * py + pz * pz);
            }
            if (!exists || metric < best_metric) { exists = true; best_metric = metric; out.face = face; out.obj = k; out.t = t; }
        }
    }
    return exists;
}

// ------------------------------------------------------------------------------------------------
// shading of one primary hit (both modes): tracer.go:53-77.  Returns the lit mask.
// ------------------------------------------------------------------------------------------------
// geometry of one primary hit in the reference's arithmetic: world hit point (object.go:109) and normal (object.go:90-95)
struct HitGeom { dv P, N; };
RT_HD HitGeom hit_geometry(const SceneView &sc, const FrameParams &fp, int i, int j, int obj, int face, float t32) {
    dv cam = dv_ld(fp.cam_pos);
    dv d = primary_dir_exact(i, j, fp);                                         // tracer.go:86
    const ObjRec &ob = sc.objs[obj];
    const MeshView &m = sc.meshes[ob.mesh];
    dv pos = dv_ld(ob.pos);
    const double *pv = m.fverts + 9 * (size_t)face;
    dv ol = dv_sub(cam, pos);                                                    // object.go:71
    TriHit h;
    const dv V1 = dv_ld_once(pv), V2 = dv_ld_once(pv + 3), V3 = dv_ld_once(pv + 6);
    if (!tri_intersection_exact(V1, V2, V3, ol, d, h)) {
        // FP32_FAST only: the fp32 winner is rejected by the fp64 test by a hair -> evaluate it anyway
        if (!tri_intersection_exact(V1, V2, V3, ol, d, h, true)) {
            h.t = (double)t32; h.ip = dv_add(ol, dv_scale(d, h.t)); h.r1 = 1.0; h.r2 = 0.0; h.r3 = 0.0;
        }
    }
    HitGeom g;
    g.P = dv_add(h.ip, pos);                                                     // object.go:109
    g.N = hit_normal_exact(m, face, h.r1, h.r2, h.r3);                           // object.go:90-95
    return g;
}

// tracer.go:61-64 : is light l visible from hit point P of object `obj`?  Ld = unit vector towards the light.
template <bool EXACT, bool DEFER = (RT_ANY_DEFER != 0)>
RT_HD bool light_visible(const SceneView &sc, const FrameParams &fp, dv P, int obj, int l, bool brute, dv &Ld, Counters *cnt) {
    const LightRec &L = sc.lights[l];
    dv Lv = dv_sub(dv_ld(L.pos), P);
    Ld = dv_norm(Lv);                                                            // tracer.go:61
    dv so = dv_add(P, dv_scale(Ld, RT_SHADOW_BIAS));                             // tracer.go:64
    double lightLen = dv_len(Lv);
    if (EXACT) {
        TraceExact s;
        bool shaded = trace_exact(sc, fp, so, Ld, brute, s, cnt);
        return !shaded || lightLen < dv_len(dv_sub(s.ip, P));                    // tracer.go:64
    } else if (fp.n_obj == 1) {
        // one object: "nearest hit, then compare with the light distance" == any hit with
        // t + bias <= lightLen  
        const ObjRec &ob = sc.objs[obj];
        RayF r = make_rayf((float)(so.x - ob.pos[0]), (float)(so.y - ob.pos[1]), (float)(so.z - ob.pos[2]), (float)Ld.x,
                           (float)Ld.y, (float)Ld.z);
        float tmax = (float)(lightLen - RT_SHADOW_BIAS);
        int f = -1;
        return !mesh_trace_fast<true, DEFER>(sc.meshes[ob.mesh], r, tmax, f, cnt, sc.trav_smem, sc.defer);
    } else {
        TraceFast s;
        bool shaded = trace_fast_nearest(sc, fp, so, Ld, false, s, cnt);
        return !shaded || lightLen < (double)s.t + RT_SHADOW_BIAS;
    }
}

RT_HD const double *hit_material(const SceneView &sc, int obj, int face) {
    const MeshView &m = sc.meshes[sc.objs[obj].mesh];
    return m.materials + 10 * (size_t)m.fmat[face];
}

// tracer.go:53-77 with the shadow queries already answered (lit_mask bit l = light l visible): lights in order
RT_HD void phong_from_mask(const SceneView &sc, const FrameParams &fp, const HitGeom &g, const double *mat, uint32_t lit_mask,
                           double col[3]) {
    dv cam = dv_ld(fp.cam_pos);
    double c[3] = { mat[0], mat[1], mat[2] };                                    // tracer.go:56
    for (int l = 0; l < fp.n_lights; l++) {
        if ((lit_mask >> l) & 1u) {
            const LightRec &L = sc.lights[l];
            dv Ld = dv_norm(dv_sub(dv_ld(L.pos), g.P));                          // tracer.go:61 (same bits as in light_visible)
            phong_light_exact(c, mat, g.N, Ld, cam, g.P, L.col);                 // tracer.go:65-72
        }
    }
    col[0] = c[0]; col[1] = c[1]; col[2] = c[2];
}

// ------------------------------------------------------------------------------------------------
// OPT-IN (RT_B200_SHADE32=1, FP32_FAST only; default off, unmeasured on the GPU so far): the whole of tracer.go:53-77 in
// fp32.  The default FAST path keeps geometry + Phong in fp64 in the reference's operation order (bit-identical colours
// wherever face and lit mask agree); fp64 is ~15 % of K4's instructions (21 divisions + 8 square roots per hit with two
// lights).  The tolerance of north_star (<= 1 LSB on >= 99.9 % of the pixels) does not need it: an fp32 colour is within
// ~1e-6 of the fp64 one, which moves uint8(255 c) by one step on ~0.05 % of the channels (tests/test_emul.py).
// ------------------------------------------------------------------------------------------------
struct fv3 { float x, y, z; };
RT_HD fv3 fv_make(float x, float y, float z) { fv3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD fv3 fv_sub(fv3 a, fv3 b) { return fv_make(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD float fv_dot(fv3 a, fv3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RT_HD fv3 fv_scale(fv3 a, float s) { return fv_make(a.x * s, a.y * s, a.z * s); }
RT_HD fv3 fv_norm(fv3 a) { return fv_scale(a, 1.0f / sqrtf(fv_dot(a, a))); }
RT_HD fv3 fv_ldd(const double *p) { return fv_make((float)p[0], (float)p[1], (float)p[2]); }
RT_HD float clamp01f(float v) { return fmaxf(0.0f, fminf(v, 1.0f)); }
// x^ns for x in [0,1]; integral exponents (every MTL file of the BASELINE configs) by square-and-multiply
RT_HD float pow_ns32(float x, float ns) {
    if (ns == 0.0f) return 1.0f;
    const float fl = floorf(ns);
    if (fl == ns && ns > 0.0f && ns <= 1024.0f) {
        float r = 1.0f, b = x;
        for (int e = (int)ns; e; e >>= 1) { if (e & 1) r *= b; b *= b; }
        return r;
    }
    return x > 0.0f ? powf(x, ns) : (ns > 0.0f ? 0.0f : 1.0f / 0.0f);
}

RT_HD uint32_t shade_hit_fast32(const SceneView &sc, const FrameParams &fp, int i, int j, int obj, int face, float t32,
                                double col[3], Counters *cnt) {
    const ObjRec &ob = sc.objs[obj];
    const MeshView &m = sc.meshes[ob.mesh];
    // the primary ray exactly as K2 shot it (fp32), object space
    const int hw = fp.W / 2, hh = fp.H / 2;
    const float phw = (float)fp.tan_half, phh = phw * (float)fp.H / (float)fp.W;
    const float si = phw * ((float)(hw - i) - 0.5f) / (float)hw, sj = phh * ((float)(hh - j) - 0.5f) / (float)hh;
    const fv3 d = fv_norm(fv_make((float)fp.fwd[0] + (float)fp.left[0] * si + (float)fp.up[0] * sj,
                                  (float)fp.fwd[1] + (float)fp.left[1] * si + (float)fp.up[1] * sj,
                                  (float)fp.fwd[2] + (float)fp.left[2] * si + (float)fp.up[2] * sj));
    const fv3 cam = fv_ldd(fp.cam_pos), pos = fv_ldd(ob.pos);
    const fv3 ol = fv_sub(cam, pos);                                                // object.go:71
    const double *pv = m.fverts + 9 * (size_t)face;
    const fv3 P1 = fv_ldd(pv), P2 = fv_ldd(pv + 3), P3 = fv_ldd(pv + 6);
    // barycentric coordinates of the hit (triangle.go:50-60) from the fp32 intersection; t32 is K2's distance
    float r1 = 1.0f, r2 = 0.0f, r3 = 0.0f, t = t32;
    {
        rt_f4 a, b, c;
        a.x = P1.x; a.y = P1.y; a.z = P1.z; a.w = 0.0f; b.x = P2.x; b.y = P2.y; b.z = P2.z; b.w = 0.0f; c.x = P3.x; c.y = P3.y; c.z = P3.z; c.w = 0.0f;
        float tt, u, v;
        if (tri_test32(a, b, c, make_rayf(ol.x, ol.y, ol.z, d.x, d.y, d.z), tt, u, v)) { t = tt; r2 = u; r3 = v; r1 = 1.0f - u - v; }
    }
    const fv3 Pl = fv_make(fmaf(d.x, t, ol.x), fmaf(d.y, t, ol.y), fmaf(d.z, t, ol.z));     // triangle.go:69
    const fv3 P = fv_make(Pl.x + pos.x, Pl.y + pos.y, Pl.z + pos.z);                         // object.go:109
    fv3 N;
    if (m.fnormals) {                                                                          // object.go:90-92
        const double *pn = m.fnormals + 9 * (size_t)face;
        const fv3 N1 = fv_ldd(pn), N2 = fv_ldd(pn + 3), N3 = fv_ldd(pn + 6);
        N = fv_norm(fv_make(N1.x * r1 + N2.x * r2 + N3.x * r3, N1.y * r1 + N2.y * r2 + N3.y * r3, N1.z * r1 + N2.z * r2 + N3.z * r3));
    } else {                                                                                   // object.go:93-95
        const fv3 e1 = fv_sub(P2, P1), e2 = fv_sub(P3, P1);
        N = fv_norm(fv_make(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x));
    }
    const double *mat = hit_material(sc, obj, face);
    float c[3] = { (float)mat[0], (float)mat[1], (float)mat[2] };                             // tracer.go:56
    const float ns = (float)mat[9];
    const fv3 V = fv_norm(fv_sub(cam, P));                                                     // tracer.go:66 (light-independent)
    uint32_t lit_mask = 0;
    for (int l = 0; l < fp.n_lights; l++) {
        RT_SYNCWARP();
        const LightRec &L = sc.lights[l];
        const fv3 Lv = fv_sub(fv_ldd(L.pos), P);
        const float lightLen = sqrtf(fv_dot(Lv, Lv));
        const fv3 Ld = fv_scale(Lv, 1.0f / lightLen);                                          // tracer.go:61
        const float bias = (float)RT_SHADOW_BIAS;
        bool lit;
        if (fp.n_obj == 1) {
            RayF r = make_rayf(fmaf(Ld.x, bias, Pl.x), fmaf(Ld.y, bias, Pl.y), fmaf(Ld.z, bias, Pl.z), Ld.x, Ld.y, Ld.z);
            float tmax = lightLen - bias;
            int f = -1;
            lit = !mesh_trace_fast<true>(m, r, tmax, f, cnt, sc.trav_smem);
        } else {
            TraceFast s;
            dv so = dv_make((double)fmaf(Ld.x, bias, P.x), (double)fmaf(Ld.y, bias, P.y), (double)fmaf(Ld.z, bias, P.z));
            bool shaded = trace_fast_nearest(sc, fp, so, dv_make((double)Ld.x, (double)Ld.y, (double)Ld.z), false, s, cnt);
            lit = !shaded || lightLen < s.t + bias;
        }
        RT_SYNCWARP();
        if (lit) {
            lit_mask |= 1u << (l & 31);
            const float ldn = fv_dot(Ld, N);
            const fv3 R = fv_sub(fv_scale(N, 2.0f * ldn), Ld);                                 // tracer.go:65
            const float diff = fmaxf(ldn, 0.0f);                                               // tracer.go:69
            const float spec = pow_ns32(fmaxf(fv_dot(R, V), 0.0f), ns);                        // tracer.go:72
            for (int k = 0; k < 3; k++) {
                const float lc = (float)L.col[k];
                c[k] = fminf(c[k] + clamp01f(diff * (float)mat[3 + k]) * lc, 1.0f);            // colour.go:38-50
                c[k] = fminf(c[k] + clamp01f(spec * (float)mat[6 + k]) * lc, 1.0f);
            }
        }
    }
    col[0] = (double)c[0]; col[1] = (double)c[1]; col[2] = (double)c[2];
    return lit_mask;
}

// fused form: geometry, one shadow query per light, Phong accumulation (both modes).  Returns the lit mask.
template <bool EXACT, bool DEFER = (RT_ANY_DEFER != 0)>
RT_HD uint32_t shade_hit(const SceneView &sc, const FrameParams &fp, int i, int j, int obj, int face, float t32,
                         bool brute, double col[3], Counters *cnt) {
    dv cam = dv_ld(fp.cam_pos);
    HitGeom g = hit_geometry(sc, fp, i, j, obj, face, t32);
    const double *mat = hit_material(sc, obj, face);
    double c[3] = { mat[0], mat[1], mat[2] };                                    // tracer.go:56
    uint32_t lit_mask = 0;
    for (int l = 0; l < fp.n_lights; l++) {
        RT_SYNCWARP();                       // every lane of the warp runs this loop (k_shade): reconverge per light
        dv Ld;
        bool lit = light_visible<EXACT, DEFER>(sc, fp, g.P, obj, l, brute, Ld, cnt);
        RT_SYNCWARP();
        if (lit) {
            lit_mask |= 1u << (l & 31);
            phong_light_exact(c, mat, g.N, Ld, cam, g.P, sc.lights[l].col);      // tracer.go:65-72
        }
    }
    col[0] = c[0]; col[1] = c[1]; col[2] = c[2];
    return lit_mask;
}

