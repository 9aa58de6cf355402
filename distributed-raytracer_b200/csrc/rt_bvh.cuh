// rt_bvh.cuh -- per-element building blocks of the on-GPU LBVH builder (K1), host+device so that
// tests/emul can replay the build on the CPU.
//
// Replaces the reference's serial R-tree construction: rtreego.NewTree(3,2,5) + Insert per face,
// /root/reference/shared/state/mesh.go:139,208.  Algorithm: Morton codes of face-box centroids,
// radix sort, Karras 2012 "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d
// Trees" hierarchy emission, bottom-up refit with SAH-guided leaf collapse, BFS re-emission into
// the 64-byte traversal nodes of rt_types.h.
#pragma once
#include "rt_types.h"
#include <math.h>
#include <string.h>

// 21 bits -> every third bit of a 63-bit word
RT_HD unsigned long long morton_spread21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8))  & 0x100f00f00f00f00full;
    v = (v | (v << 4))  & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2))  & 0x1249249249249249ull;
    return v;
}
// x,y,z in [0,1]
RT_HD unsigned long long morton63(double x, double y, double z) {
    const double s = 2097152.0;   // 2^21
    double fx = fmin(fmax(x * s, 0.0), s - 1.0), fy = fmin(fmax(y * s, 0.0), s - 1.0), fz = fmin(fmax(z * s, 0.0), s - 1.0);
    return (morton_spread21((unsigned long long)fx) << 2) | (morton_spread21((unsigned long long)fy) << 1) |
           morton_spread21((unsigned long long)fz);
}

RT_HD int rt_clz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
RT_HD int rt_clz32(unsigned int v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// Karras delta: length of the common prefix of keys i and j (index used as a tie-breaker), -1 outside
RT_HD int karras_delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + rt_clz32((unsigned int)i ^ (unsigned int)j);
    return rt_clz64(a ^ b);
}

// Child encoding inside the builder: >= 0 internal node index, < 0 : ~leafIndex (sorted position)
struct KarrasNode { int left, right, lo, hi; };

// Internal node i of the binary radix tree over n sorted keys (n >= 2, 0 <= i <= n-2)
RT_HD KarrasNode karras_node(const unsigned long long *keys, int n, int i) {
    int dl = karras_delta(keys, n, i, i - 1), dr = karras_delta(keys, n, i, i + 1);
    int d = (dr - dl) < 0 ? -1 : 1;
    int dmin = d > 0 ? dl : dr;
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karras_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    KarrasNode k;
    k.lo = i < j ? i : j;
    k.hi = i < j ? j : i;
    k.left = (k.lo == gamma) ? ~gamma : gamma;
    k.right = (k.hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    return k;
}

// round a double outward to float (lo: toward -inf, hi: toward +inf)
RT_HD float f32_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -HUGE_VALF);
    return f;
}
RT_HD float f32_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, HUGE_VALF);
    return f;
}

struct BoxF { float lo[3], hi[3]; };

RT_HD BoxF box_union(const BoxF &a, const BoxF &b) {
    BoxF r;
    for (int k = 0; k < 3; k++) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}
RT_HD float box_half_area(const BoxF &b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

#define RT_SAH_CI 1.2f   // cost of one node visit relative to one triangle test
#define RT_SAH_CT 1.0f

// ------------------------------------------------------------------------------------------------
// Kernel bodies, one call per (emulated) thread.  bvh_build.cu wraps each in a __global__ kernel;
// tests/emul runs them in plain loops on the CPU.
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define RT_ATOMIC_INC(p) atomicAdd((p), 1u)
#define RT_FENCE() __threadfence()
#define RT_LDCG_F(p) __ldcg(p)
#else
#define RT_ATOMIC_INC(p) ((*(p))++)
#define RT_FENCE() do { } while (0)
#define RT_LDCG_F(p) (*(p))
#endif

struct BuildParams {
    float lo[3], hi[3];   // mesh bounds
    float pad;            // absolute padding of every leaf box (2^-21 of the mesh extent/magnitude)
};

// mesh bounds (host side of the step; the device reduces with ordered-int atomics) -> BuildParams + padded mesh box
RT_HD BuildParams make_build_params(const float lo[3], const float hi[3], float mlo[3], float mhi[3]) {
    BuildParams bp;
    float ext = 0.f, mag = 0.f;
    for (int k = 0; k < 3; k++) {
        bp.lo[k] = lo[k]; bp.hi[k] = hi[k];
        ext = fmaxf(ext, hi[k] - lo[k]);
        mag = fmaxf(mag, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
    }
    // 2^-21 of the mesh size: far above any fp64 rounding of a hit point, far below a triangle
    bp.pad = fmaxf(fmaxf(ext, mag) * 4.76837158e-7f, 1e-30f);
    for (int k = 0; k < 3; k++) { mlo[k] = f32_down((double)lo[k] - 2.0 * bp.pad); mhi[k] = f32_up((double)hi[k] + 2.0 * bp.pad); }
    return bp;
}

// per face: padded fp32 leaf box + Morton key of the box centroid
RT_HD void leaf_morton_thread(uint32_t f, const double *fverts, const BuildParams &bp, BoxF *leafbox,
                              unsigned long long *keys, uint32_t *vals) {
    const double *p = fverts + 9 * (size_t)f;
    BoxF bx;
    double cn[3];
    for (int k = 0; k < 3; k++) {
        double a = p[k], b = p[3 + k], c = p[6 + k];
        double lo = fmin(a, fmin(b, c)), hi = fmax(a, fmax(b, c));
        bx.lo[k] = f32_down(lo - (double)bp.pad);
        bx.hi[k] = f32_up(hi + (double)bp.pad);
        double ext = (double)bp.hi[k] - (double)bp.lo[k];
        cn[k] = ext > 0.0 ? (0.5 * (lo + hi) - (double)bp.lo[k]) / ext : 0.0;
    }
    leafbox[f] = bx;
    keys[f] = morton63(cn[0], cn[1], cn[2]);
    vals[f] = f;
}

RT_HD void karras_thread(int i, const unsigned long long *keys, int n, KarrasNode *kn, int *parent_int, int *parent_leaf) {
    KarrasNode k = karras_node(keys, n, i);
    kn[i] = k;
    if (k.left >= 0) parent_int[k.left] = i; else parent_leaf[~k.left] = i;
    if (k.right >= 0) parent_int[k.right] = i; else parent_leaf[~k.right] = i;
    if (i == 0) parent_int[0] = -1;
}

RT_HD BoxF ld_box_cg(const BoxF *p) {
    BoxF b;
    const float *q = reinterpret_cast<const float *>(p);
    for (int k = 0; k < 3; k++) { b.lo[k] = RT_LDCG_F(q + k); b.hi[k] = RT_LDCG_F(q + 3 + k); }
    return b;
}

// One thread per leaf walks up; the second thread to reach a node owns it: union of the child boxes,
// SAH cost, and the decision to collapse the subtree (<= max_leaf faces) into one leaf.
RT_HD void refit_sah_thread(int i, const uint32_t *vals, const BoxF *leafbox, const KarrasNode *kn, const int *parent_int,
                            const int *parent_leaf, BoxF *ibox, float *icost, unsigned char *collapsed, unsigned int *flags,
                            int max_leaf) {
    int p = parent_leaf[i];
    while (p >= 0) {
        RT_FENCE();
        unsigned int old = RT_ATOMIC_INC(&flags[p]);
        if (old == 0) return;                       // sibling subtree not finished yet
        RT_FENCE();
        KarrasNode k = kn[p];
        BoxF bl, br;
        float cl, cr;
        if (k.left >= 0) { bl = ld_box_cg(&ibox[k.left]); cl = RT_LDCG_F(&icost[k.left]); }
        else { bl = leafbox[vals[~k.left]]; cl = RT_SAH_CT * box_half_area(bl); }
        if (k.right >= 0) { br = ld_box_cg(&ibox[k.right]); cr = RT_LDCG_F(&icost[k.right]); }
        else { br = leafbox[vals[~k.right]]; cr = RT_SAH_CT * box_half_area(br); }
        BoxF b = box_union(bl, br);
        float area = box_half_area(b);
        int count = k.hi - k.lo + 1;
        float c_int = RT_SAH_CI * area + cl + cr;
        float c_leaf = RT_SAH_CT * area * (float)count;
        bool col = count <= max_leaf && c_leaf <= c_int;
        ibox[p] = b;
        icost[p] = col ? c_leaf : c_int;
        collapsed[p] = col ? 1 : 0;
        p = parent_int[p];
    }
}

// 0: leaf (single face), 1: collapsed subtree (leaf), 2: internal
RT_HD int child_class(int c, const unsigned char *collapsed) {
    if (c < 0) return 0;
    return collapsed[c] ? 1 : 2;
}

// Children of the wide node rooted at Karras node `node`: start from its two children and keep replacing the
// expandable child (internal, not collapsed) of largest surface area by its own two children until RT_W slots are
// used.  out[] holds builder child codes (>= 0 internal Karras index, < 0 ~leaf).  Deterministic.
RT_HD int expand_children(int node, const KarrasNode *kn, const unsigned char *collapsed, const BoxF *ibox, int out[RT_W]) {
    out[0] = kn[node].left; out[1] = kn[node].right;
    int n = 2;
    while (n < RT_W) {
        int best = -1;
        float best_a = -1.0f;
        for (int i = 0; i < n; i++)
            if (child_class(out[i], collapsed) == 2) {
                float a = box_half_area(ibox[out[i]]);
                if (a > best_a) { best_a = a; best = i; }
            }
        if (best < 0) break;
        int c = out[best];
        out[best] = kn[c].left;
        out[n++] = kn[c].right;
    }
    return n;
}

RT_HD void emit_count_thread(int p, const int *frontier, const KarrasNode *kn, const unsigned char *collapsed, const BoxF *ibox,
                             int *cnt) {
    int ch[RT_W];
    int n = expand_children(frontier[p], kn, collapsed, ibox, ch);
    int c = 0;
    for (int i = 0; i < n; i++) c += child_class(ch[i], collapsed) == 2;
    cnt[p] = c;
}

// writes traversal node (level_base + p); internal children go to the next BFS level. Returns #leaf children.
RT_HD unsigned int emit_write_thread(int p, const int *frontier, int fsize, int level_base, const KarrasNode *kn,
                                     const unsigned char *collapsed, const int *off, const uint32_t *vals, const BoxF *leafbox,
                                     const BoxF *ibox, RtNode *nodes, int *next_frontier) {
    int ch[RT_W];
    int n = expand_children(frontier[p], kn, collapsed, ibox, ch);
    int o = off[p];
    int next_base = level_base + fsize;
    unsigned int nleaf = 0;
    RtNode nd;
    for (int s = 0; s < RT_W; s++) {
        BoxF bx;
        int ref;
        if (s >= n) {
            for (int k = 0; k < 3; k++) { bx.lo[k] = 0.f; bx.hi[k] = 0.f; }
            ref = RT_EMPTY;
        } else {
            int c = ch[s];
            int cls = child_class(c, collapsed);
            if (cls == 0) {
                bx = leafbox[vals[~c]];
                ref = rt_leaf_ref((uint32_t)~c, 1u);
                nleaf++;
            } else if (cls == 1) {
                bx = ibox[c];
                KarrasNode ck = kn[c];
                ref = rt_leaf_ref((uint32_t)ck.lo, (uint32_t)(ck.hi - ck.lo + 1));
                nleaf++;
            } else {
                bx = ibox[c];
                ref = next_base + o;
                next_frontier[o] = c;
                o++;
            }
        }
        nd.lox[s] = bx.lo[0]; nd.hix[s] = bx.hi[0]; nd.loy[s] = bx.lo[1]; nd.hiy[s] = bx.hi[1];
        nd.loz[s] = bx.lo[2]; nd.hiz[s] = bx.hi[2]; nd.ref[s] = ref; nd.pad[s] = 0;
    }
    nodes[level_base + p] = nd;
    return nleaf;
}

// grid of the quantised nodes from the padded mesh bounds (rt_types.h: RtNodeQ): two cells of margin on either side
RT_HD void qgrid_from_bounds(const float lo[3], const float hi[3], float qorg[3], float qstep[3]) {
    for (int k = 0; k < 3; k++) {
        float st = (hi[k] - lo[k]) / RT_QGRID_CELLS;
        if (!(st > 1e-30f)) st = 1e-30f;
        qstep[k] = st;
        qorg[k] = lo[k] - 2.0f * st;
    }
}
RT_HD uint16_t qgrid_down(float v, float org, float st) {
    double q = floor(((double)v - (double)org) / (double)st) - 1.0;
    return (uint16_t)(q < 0.0 ? 0.0 : (q > 65535.0 ? 65535.0 : q));
}
RT_HD uint16_t qgrid_up(float v, float org, float st) {
    double q = ceil(((double)v - (double)org) / (double)st) + 1.0;
    return (uint16_t)(q < 0.0 ? 0.0 : (q > 65535.0 ? 65535.0 : q));
}
// node i of a float hierarchy -> its 64-byte quantised form (every box grows by at least one cell on every side)
RT_HD void quantise_node_thread(uint32_t i, const RtNode *nodes, RtNodeQ *out, const float qorg[3], const float qstep[3]) {
    const RtNode nd = nodes[i];
    RtNodeQ q;
    for (int s = 0; s < 4; s++) {
        q.x[s] = qgrid_down(nd.lox[s], qorg[0], qstep[0]); q.x[4 + s] = qgrid_up(nd.hix[s], qorg[0], qstep[0]);
        q.y[s] = qgrid_down(nd.loy[s], qorg[1], qstep[1]); q.y[4 + s] = qgrid_up(nd.hiy[s], qorg[1], qstep[1]);
        q.z[s] = qgrid_down(nd.loz[s], qorg[2], qstep[2]); q.z[4 + s] = qgrid_up(nd.hiz[s], qorg[2], qstep[2]);
        q.ref[s] = nd.ref[s];
    }
    out[i] = q;
}

RT_HD float rt_int_as_float(int v) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// triangle records in leaf (sorted) order
RT_HD void gather_tri_thread(uint32_t i, const double *fverts, const uint32_t *vals, rt_f4 *tri32, double *tri64) {
    uint32_t f = vals[i];
    const double *p = fverts + 9 * (size_t)f;
    double v[9];
    for (int k = 0; k < 9; k++) v[k] = p[k];
    if (tri32) {
        rt_f4 a, b, c;
        a.x = (float)v[0]; a.y = (float)v[1]; a.z = (float)v[2]; a.w = rt_int_as_float((int)f);
        b.x = (float)v[3]; b.y = (float)v[4]; b.z = (float)v[5]; b.w = 0.0f;
        c.x = (float)v[6]; c.y = (float)v[7]; c.z = (float)v[8]; c.w = 0.0f;
        rt_f4 *q = tri32 + RT_TRI32_F4 * (size_t)i;
        q[0] = a; q[1] = b; q[2] = c;
        if (RT_TRI32_F4 > 3) { rt_f4 z; z.x = z.y = z.z = z.w = 0.0f; q[3] = z; }
    }
    if (tri64) {
        double *q = tri64 + 10 * (size_t)i;
        for (int k = 0; k < 9; k++) q[k] = v[k];
        q[9] = (double)f;
    }
}
