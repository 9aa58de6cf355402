// rt_exact.cuh -- float64 arithmetic in the reference's operation order (FP64_EXACT mode and the
// shading of both modes).  Every +,-,*,/ and sqrt is a separately rounded IEEE binary64 operation:
// on the device the __d*_rn intrinsics are used, which nvcc never contracts into FMA regardless of
// -fmad; on the host (tests/emul only) the file must be compiled with -ffp-contract=off.
//
// Reference (paths under /root/reference/): shared/geom/vector.go, shared/geom/triangle.go,
// shared/colour/colour.go, worker/shared/tracer/tracer.go, Go stdlib math.Pow / math.Max / math.Min.
#pragma once
#include "rt_types.h"
#include <math.h>

#if defined(__CUDA_ARCH__)
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DMUL(a, b) __dmul_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#define DSQRT(a) __dsqrt_rn((a))
#else
#define DADD(a, b) ((a) + (b))
#define DSUB(a, b) ((a) - (b))
#define DMUL(a, b) ((a) * (b))
#define DDIV(a, b) ((a) / (b))
#define DSQRT(a) sqrt((a))
#endif

struct dv { double x, y, z; };

RT_HD dv dv_make(double x, double y, double z) { dv r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD dv dv_ld(const double *p) { return dv_make(p[0], p[1], p[2]); }
// per-hit records (float64 vertices / normals of the winning face) are read once per pixel: streaming loads keep them from evicting
// BVH nodes out of L1 (RT_HIT_STREAM, compile-time experiment)
#ifndef RT_HIT_STREAM
#define RT_HIT_STREAM 0
#endif
RT_HD dv dv_ld_once(const double *p) {
#if defined(__CUDA_ARCH__) && RT_HIT_STREAM
    return dv_make(__ldcs(p), __ldcs(p + 1), __ldcs(p + 2));
#else
    return dv_make(p[0], p[1], p[2]);
#endif
}
RT_HD dv dv_add(dv a, dv b) { return dv_make(DADD(a.x, b.x), DADD(a.y, b.y), DADD(a.z, b.z)); }              // vector.go:14
RT_HD dv dv_sub(dv a, dv b) { return dv_make(DSUB(a.x, b.x), DSUB(a.y, b.y), DSUB(a.z, b.z)); }              // vector.go:19
RT_HD dv dv_scale(dv a, double s) { return dv_make(DMUL(s, a.x), DMUL(s, a.y), DMUL(s, a.z)); }              // vector.go:24
RT_HD double dv_dot(dv a, dv b) { return DADD(DADD(DMUL(a.x, b.x), DMUL(a.y, b.y)), DMUL(a.z, b.z)); }       // vector.go:29
RT_HD dv dv_cross(dv a, dv b) {                                                                              // vector.go:34
    return dv_make(DSUB(DMUL(a.y, b.z), DMUL(a.z, b.y)), DSUB(DMUL(a.z, b.x), DMUL(a.x, b.z)),
                   DSUB(DMUL(a.x, b.y), DMUL(a.y, b.x)));
}
RT_HD double dv_len(dv a) { return DSQRT(DADD(DADD(DMUL(a.x, a.x), DMUL(a.y, a.y)), DMUL(a.z, a.z))); }      // vector.go:56
RT_HD dv dv_norm(dv a) {                                                                                     // vector.go:50
    double mag = dv_len(a);
    return dv_make(DDIV(a.x, mag), DDIV(a.y, mag), DDIV(a.z, mag));
}

// Go math.Max / math.Min (Go stdlib math/dim.go): Max(x,+Inf)=+Inf even against NaN, otherwise NaN-propagating,
// Max(+0,-0)=+0, Min(+0,-0)=-0.  Written so that the common case (ordered, not both zero) is one compare + select;
// the rare cases sit behind two cheap tests (tests/test_emul.py checks every special case against the oracle).
RT_HD double go_max(double x, double y) {
    double r = x > y ? x : y;
    if (!(x == x && y == y)) r = (x > 1.7976931348623157e308 || y > 1.7976931348623157e308) ? HUGE_VAL : x + y;
    else if (x == 0 && y == 0) r = signbit(x) ? y : x;
    return r;
}
RT_HD double go_min(double x, double y) {
    double r = x < y ? x : y;
    if (!(x == x && y == y)) r = (x < -1.7976931348623157e308 || y < -1.7976931348623157e308) ? -HUGE_VAL : x + y;
    else if (x == 0 && y == 0) r = signbit(x) ? x : y;
    return r;
}

// ------------------------------------------------------------------------------------------------
// The reference's object-level candidate test (tracer.go:32): a ray is tried against an object only if box.go's predicate
// holds for the object's bounds (object.go:31-59).  The predicate is NOT conservative: it looks for a side whose plane the ray
// meets inside the side's rectangle, and a ray that leaves the box exactly through an edge can fail that test on both
// adjacent sides by one rounding -- the reference then never looks at the object (seen for shadow rays in grid-aligned stress
// scenes, tools/fuzz_degenerate_faces_cpu.py; never in a BASELINE config).  FP64_EXACT applies the same predicate to the
// same rectangle, operation for operation.  (The rectangles INSIDE rtreego's face tree, object.go:76, depend on that
// library's insertion order and cannot be mirrored: DESIGN.md s4.)
// ------------------------------------------------------------------------------------------------
#define RT_BOUND_EPSILON 1e-4      // boundEpsilon, shared/state/object.go / mesh.go

// box.go:29-69, operation for operation (the dot products with the axis normals keep their multiplications by 0 and +-1)
RT_HD bool box_intersect_exact(dv bmin, dv bmax, dv o, dv d) {
    for (int k = 0; k < 6; k++) {                                                       // boxNormals, box.go:7-11
        const dv n = dv_make(k == 0 ? 1.0 : (k == 1 ? -1.0 : 0.0), k == 2 ? 1.0 : (k == 3 ? -1.0 : 0.0), k == 4 ? 1.0 : (k == 5 ? -1.0 : 0.0));
        const double dn = dv_dot(d, n);                                                 // :32
        if (dn != 0.0) {
            const dv sp = (dv_dot(n, dv_make(1.0, 1.0, 1.0)) < 0) ? bmin : bmax;        // :35-39
            const double dirScale = DDIV(dv_dot(dv_sub(sp, o), n), dn);                 // :42
            if (dirScale >= 0.0) {                                                      // :45
                const dv p = dv_add(o, dv_scale(d, dirScale));                          // :47
                if (n.x != 0.0) { if ((bmin.y <= p.y && p.y <= bmax.y) && (bmin.z <= p.z && p.z <= bmax.z)) return true; }          // :50-53
                else if (n.y != 0.0) { if ((bmin.x <= p.x && p.x <= bmax.x) && (bmin.z <= p.z && p.z <= bmax.z)) return true; }     // :54-57
                else if (n.z != 0.0) { if ((bmin.x <= p.x && p.x <= bmax.x) && (bmin.y <= p.y && p.y <= bmax.y)) return true; }     // :58-61
            }
        }
    }
    return false;
}

// Object.Bounds() (object.go:31-59) + geom.NewBox (box.go:21-26) from the bounds of ALL vertices of the mesh: fl(P + v) is
// monotone in v, so the min / max over the vertices of fl(P.x + v.x) is fl(P.x + min / max v.x).  rtreego.NewRect(p, len) keeps
// q = p + len; NewBox returns MinCorner = p, MaxCorner = p + (q - p).  No vertices: vlo = +Inf, vhi = -Inf => the box is P.
RT_HD void object_box_exact(dv pos, dv vlo, dv vhi, dv &bmin, dv &bmax) {
    const dv lo = dv_make(go_min(pos.x, DADD(pos.x, vlo.x)), go_min(pos.y, DADD(pos.y, vlo.y)), go_min(pos.z, DADD(pos.z, vlo.z)));
    const dv hi = dv_make(go_max(pos.x, DADD(pos.x, vhi.x)), go_max(pos.y, DADD(pos.y, vhi.y)), go_max(pos.z, DADD(pos.z, vhi.z)));
    const dv len = dv_make(go_max(DSUB(hi.x, lo.x), RT_BOUND_EPSILON), go_max(DSUB(hi.y, lo.y), RT_BOUND_EPSILON), go_max(DSUB(hi.z, lo.z), RT_BOUND_EPSILON));
    const dv q = dv_add(lo, len);
    bmin = lo;
    bmax = dv_add(lo, dv_sub(q, lo));
}

RT_HD bool go_is_odd_int(double x) {
    double xi;
    double xf = modf(x, &xi);
    return xf == 0 && (((long long)xi) & 1) == 1;
}

// Go math.Pow (math/pow.go).  Integer exponents (every BASELINE config: Ns 10 / Ns 0) take the
// Frexp / square-and-multiply / Ldexp path: multiplications only => bit-reproducible on the GPU.
// A fractional exponent uses exp(yf*log(x)) (Go's Exp/Log are arch-specific: parity unpinned).
RT_HD double go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (x != x || y != y) return x + y;
    if (x == 0) {
        if (y < 0) { if (go_is_odd_int(y)) return copysign(HUGE_VAL, x); return HUGE_VAL; }
        if (y > 0) { if (go_is_odd_int(y)) return x; return 0; }
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return HUGE_VAL;
    }
    if (isinf(x)) {
        if (x < 0) { x = DDIV(1.0, x); y = -y; if (y < 0) { return go_is_odd_int(y) ? copysign(HUGE_VAL, x) : HUGE_VAL; } return go_is_odd_int(y) ? x : 0; }
        if (y < 0) return 0;
        if (y > 0) return HUGE_VAL;
    }
    if (y == 0.5) return DSQRT(x);
    if (y == -0.5) return DDIV(1.0, DSQRT(x));

    double absy = y;
    bool flip = false;
    if (absy < 0) { absy = -absy; flip = true; }
    double yi;
    double yf = modf(absy, &yi);
    if (yf != 0 && x < 0) return (x - x) / (x - x);   // NaN (x finite here)
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return HUGE_VAL;
    }
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf = DSUB(yf, 1.0); yi = DADD(yi, 1.0); }
        a1 = exp(DMUL(yf, log(x)));
    }
    int xe;
    double x1 = frexp(x, &xe);
    for (long long i = (long long)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 = DMUL(a1, x1); ae += xe; }
        x1 = DMUL(x1, x1);
        xe <<= 1;
        if (x1 < .5) { x1 = DADD(x1, x1); xe--; }
    }
    if (flip) { a1 = DDIV(1.0, a1); ae = -ae; }
    return ldexp(a1, ae);
}

// triangle.go:37-77.  force=true skips the acceptance tests (used by FP32_FAST to evaluate the
// fp32 winner's hit point in fp64 even if the fp64 test would reject it by a hair).
struct TriHit { dv ip; double r1, r2, r3, t; };
RT_HD bool tri_intersection_exact(dv P1, dv P2, dv P3, dv o, dv d, TriHit &h, bool force = false) {
    dv p1p2 = dv_sub(P2, P1), p1p3 = dv_sub(P3, P1), nd = dv_scale(d, -1.0);          // :38
    dv c = dv_cross(p1p3, nd);
    double incidence = dv_dot(p1p2, c);                                              // :42
    if (incidence != 0.0) {                                                          // :46
        dv p1or = dv_sub(o, P1);                                                     // :47
        double r2 = DDIV(dv_dot(p1or, c), incidence);                                // :50 (same bits as recomputing the cross)
        if (force || (0.0 <= r2 && r2 <= 1.0)) {                                     // :53
            double r3 = DDIV(dv_dot(p1p2, dv_cross(p1or, nd)), incidence);           // :55
            double s = DADD(r2, r3);
            if (force || (0.0 <= s && s <= 1.0)) {                                   // :58
                double r1 = DSUB(DSUB(1.0, r2), r3);                                 // :60
                if (force || (r1 >= 0.0 && r2 >= 0.0 && r3 >= 0.0)) {                // :63
                    double t = DDIV(dv_dot(p1p2, dv_cross(p1p3, p1or)), incidence);  // :65
                    if (force || t >= 0.0) {                                         // :68
                        h.ip = dv_add(o, dv_scale(d, t));                            // :69
                        h.r1 = r1; h.r2 = r2; h.r3 = r3; h.t = t;
                        return true;
                    }
                }
            }
        }
    }
    return false;
}

// tracer.go:15-22 + :86 : primary ray direction of pixel (i,j); tan(Fov/2) is a frame constant.
RT_HD dv primary_dir_exact(int i, int j, const FrameParams &fp) {
    int halfWidth = fp.W / 2, halfHeight = fp.H / 2;                                                   // :16
    double phw = fp.tan_half;                                                                          // :17
    double phh = DDIV(DMUL(phw, (double)fp.H), (double)fp.W);                                          // :18
    double si = DDIV(DMUL(phw, DSUB((double)(halfWidth - i), 0.5)), (double)halfWidth);                // :19
    double sj = DDIV(DMUL(phh, DSUB((double)(halfHeight - j), 0.5)), (double)halfHeight);              // :20
    dv pos = dv_ld(fp.cam_pos);
    dv pt = dv_add(dv_add(dv_add(pos, dv_ld(fp.fwd)), dv_scale(dv_ld(fp.left), si)), dv_scale(dv_ld(fp.up), sj));  // :21
    return dv_norm(dv_sub(pt, pos));                                                                   // :86
}

// triangle.go:29-31 (InterpNormal) / :24-26 (Normal), selected as in object.go:90-95
RT_HD dv hit_normal_exact(const MeshView &m, int face, double r1, double r2, double r3) {
    if (m.fnormals) {
        const double *n = m.fnormals + 9 * (size_t)face;
        return dv_norm(dv_add(dv_add(dv_scale(dv_ld_once(n), r1), dv_scale(dv_ld_once(n + 3), r2)), dv_scale(dv_ld_once(n + 6), r3)));
    }
    const double *p = m.fverts + 9 * (size_t)face;
    dv P1 = dv_ld_once(p);
    return dv_norm(dv_cross(dv_sub(dv_ld_once(p + 3), P1), dv_sub(dv_ld_once(p + 6), P1)));
}

// colour.go:38-50
RT_HD double col_add(double a, double b) { return go_min(DADD(a, b), 1.0); }
RT_HD double col_scale(double a, double s) { return go_max(0.0, go_min(DMUL(s, a), 1.0)); }

// tracer.go:65-72 : one lit light's contribution folded into c[3]
RT_HD void phong_light_exact(double c[3], const double *mat10, dv N, dv lightDir, dv camPos, dv P, const double *lcol) {
    dv reflectDir = dv_sub(dv_scale(N, DMUL(2.0, dv_dot(lightDir, N))), lightDir);          // :65
    dv camDir = dv_norm(dv_sub(camPos, P));                                                 // :66
    double diff = go_max(dv_dot(lightDir, N), 0.0);                                         // :69
    double spec = go_pow(go_max(dv_dot(reflectDir, camDir), 0.0), mat10[9]);                // :72
    for (int k = 0; k < 3; k++) c[k] = col_add(c[k], DMUL(col_scale(mat10[3 + k], diff), lcol[k]));   // :69
    for (int k = 0; k < 3; k++) c[k] = col_add(c[k], DMUL(col_scale(mat10[6 + k], spec), lcol[k]));   // :72
}

// colour.go:59-61 RGB()  /  colour.go:54-56 RGBA() followed by go-sdl2's >>8
RT_HD unsigned char quantise_exact(double c, bool sdl) {
    if (!(c == c)) return 0;
    if (sdl) return (unsigned char)(((unsigned int)DMUL(65535.0, c)) >> 8);
    return (unsigned char)(int)DMUL(255.0, c);
}
