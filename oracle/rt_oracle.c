/*
 * rt_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See rt_oracle.h.
 *
 * PARITY UNPINNED (no reference tests / golden vectors exist; reference cannot run here).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -lpthread  (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: Go on amd64 never fuses a*b+c, every
 * operation below is a separately rounded IEEE binary64 op in the order the Go source
 * writes it.
 *
 * Every function cites the reference file:line it restates (paths under /root/reference/).
 */
#define _GNU_SOURCE
#include "rt_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

/* ============================ shared/geom/vector.go ======================== */
typedef struct { double x, y, z; } vec;

static inline vec v_add(vec a, vec b) { vec r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }   /* vector.go:14 */
static inline vec v_sub(vec a, vec b) { vec r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }   /* vector.go:19 */
static inline vec v_scale(vec a, double s) { vec r = { s * a.x, s * a.y, s * a.z }; return r; }   /* vector.go:24 (scalar on the left) */
static inline double v_dot(vec a, vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            /* vector.go:29 */
static inline vec v_cross(vec a, vec b) {                                                         /* vector.go:34 */
    vec r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
static inline int v_zero(vec a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }                /* vector.go:45 */
static inline vec v_norm(vec a) {                                                                 /* vector.go:50 */
    double mag = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    vec r = { a.x / mag, a.y / mag, a.z / mag };
    return r;
}
static inline double v_len(vec a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }            /* vector.go:56 */
static inline vec v_ld(const double *p) { vec r = { p[0], p[1], p[2] }; return r; }
static inline void v_st(double *p, vec a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

/* Go math.Max / math.Min (NaN-propagating, signed-zero aware).  Go stdlib math/dim.go. */
static inline double go_max(double x, double y) {
    if (isinf(x) && x > 0) return x;
    if (isinf(y) && y > 0) return y;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) { return signbit(x) ? y : x; }
    return x > y ? x : y;
}
static inline double go_min(double x, double y) {
    if (isinf(x) && x < 0) return x;
    if (isinf(y) && y < 0) return y;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) { return signbit(x) ? x : y; }
    return x < y ? x : y;
}

double orc_go_max(double x, double y) { return go_max(x, y); }
double orc_go_min(double x, double y) { return go_min(x, y); }

/* ============================ Go stdlib math ================================ */
/* math/tan.go (pure Go Cephes routine; no amd64 assembly).  Only the x < 2^29 branch
 * (no Payne-Hanek reduction) is restated; larger |x| falls back to libm. */
double orc_go_tan(double x) {
    static const double PI4A = 7.85398125648498535156E-1;  /* 0x3fe921fb40000000 */
    static const double PI4B = 3.77489470793079817668E-8;  /* 0x3e64442d00000000 */
    static const double PI4C = 2.69515142907905952645E-15; /* 0x3ce8469898cc5170 */
    static const double P[3] = { -1.30936939181383777646E4, 1.15351664838587416140E6, -1.79565251976484877988E7 };
    static const double Q[5] = { 1.0, 1.36812963470692954678E4, -1.32089234440210967447E6,
                                 2.50083801823357915839E7, -5.38695755929454629881E7 };
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    if (x >= 536870912.0 /* reduceThreshold = 1<<29 */) { double t = tan(x); return sign ? -t : t; }
    uint64_t j = (uint64_t)(x * 1.27323954473516268615107010698011489627567716592365 /* Go constant 4/Pi, exact-rational then rounded */);
    double y = (double)j;
    if (j & 1) { j++; y++; }
    double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
    double zz = z * z;
    if (zz > 1e-14) {
        y = z + z * (zz * (((P[0] * zz) + P[1]) * zz + P[2]) / ((((zz + Q[1]) * zz + Q[2]) * zz + Q[3]) * zz + Q[4]));
    } else {
        y = z;
    }
    if (j & 2) y = -1 / y;
    if (sign) y = -y;
    return y;
}

static int is_odd_int(double x) {
    double xi;
    double xf = modf(x, &xi);
    return xf == 0 && ((int64_t)xi & 1) == 1;
}

/* math/pow.go.  The integer-exponent part (Frexp / square-and-multiply / Ldexp) is exact
 * restatement; the fractional part uses libm exp/log (Go's Exp/Log are arch-specific =>
 * unpinned; all BASELINE configs keep Ns integral). */
double orc_go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) { if (is_odd_int(y)) return copysign(INFINITY, x); return INFINITY; }
        if (y > 0) { if (is_odd_int(y)) return x; return 0; }
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (isinf(x)) {
        if (x < 0) return orc_go_pow(1 / x, -y);
        if (y < 0) return 0;
        if (y > 0) return INFINITY;
    }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);

    double absy = y;
    int flip = 0;
    if (absy < 0) { absy = -absy; flip = 1; }
    double yi;
    double yf = modf(absy, &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = exp(yf * log(x));
    }
    int xe;
    double x1 = frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (flip) { a1 = 1 / a1; ae = -ae; }
    return ldexp(a1, ae);
}

/* ============================ shared/geom/triangle.go ======================= */
/* triangle.go:37-77 */
static inline int tri_intersection(vec P1, vec P2, vec P3, vec o, vec d, vec *ip, double *r1o, double *r2o,
                                   double *r3o, double *to) {
    vec p1p2 = v_sub(P2, P1), p1p3 = v_sub(P3, P1), nd = v_scale(d, -1);   /* :38 */
    double incidence = v_dot(p1p2, v_cross(p1p3, nd));                     /* :42 */
    if (incidence != 0.0) {                                                /* :46 */
        vec p1or = v_sub(o, P1);                                           /* :47 */
        double r2 = v_dot(p1or, v_cross(p1p3, nd)) / incidence;            /* :50 */
        if (0.0 <= r2 && r2 <= 1.0) {                                      /* :53 */
            double r3 = v_dot(p1p2, v_cross(p1or, nd)) / incidence;        /* :55 */
            if (0.0 <= r2 + r3 && r2 + r3 <= 1.0) {                        /* :58 */
                double r1 = 1.0 - r2 - r3;                                 /* :60 */
                if (r1 >= 0.0 && r2 >= 0.0 && r3 >= 0.0) {                 /* :63 */
                    double dirScale = v_dot(p1p2, v_cross(p1p3, p1or)) / incidence; /* :65 */
                    if (dirScale >= 0.0) {                                 /* :68 */
                        *ip = v_add(o, v_scale(d, dirScale));              /* :69 */
                        *r1o = r1; *r2o = r2; *r3o = r3; *to = dirScale;
                        return 1;
                    }
                }
            }
        }
    }
    return 0;
}

int orc_triangle_intersection(const double *tri, const double *o, const double *d, double *ip, double *bary,
                              double *t_out) {
    vec p; double r1, r2, r3, t;
    int hit = tri_intersection(v_ld(tri), v_ld(tri + 3), v_ld(tri + 6), v_ld(o), v_ld(d), &p, &r1, &r2, &r3, &t);
    if (hit) { v_st(ip, p); bary[0] = r1; bary[1] = r2; bary[2] = r3; *t_out = t; }
    return hit;
}

/* ============================ shared/geom/box.go ============================ */
/* box.go:29-69.  boxNormals order :7-11: +x,-x,+y,-y,+z,-z.  For normal n: the plane point is
 * MinCorner when n.(1,1,1) < 0 else MaxCorner (:33-38); dirScale = (sPoint-o).n / d.n (:40). */
static inline int box_intersect(vec bmin, vec bmax, vec o, vec d) {
    static const vec N[6] = { {1,0,0}, {-1,0,0}, {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1} };
    for (int k = 0; k < 6; k++) {
        vec n = N[k];
        double dn = v_dot(d, n);
        if (dn != 0.0) {
            vec ones = { 1, 1, 1 };
            vec sp = (v_dot(n, ones) < 0) ? bmin : bmax;
            double dirScale = v_dot(v_sub(sp, o), n) / dn;
            if (dirScale >= 0.0) {
                vec p = v_add(o, v_scale(d, dirScale));
                if (n.x != 0.0) {
                    if ((bmin.y <= p.y && p.y <= bmax.y) && (bmin.z <= p.z && p.z <= bmax.z)) return 1;
                } else if (n.y != 0.0) {
                    if ((bmin.x <= p.x && p.x <= bmax.x) && (bmin.z <= p.z && p.z <= bmax.z)) return 1;
                } else if (n.z != 0.0) {
                    if ((bmin.x <= p.x && p.x <= bmax.x) && (bmin.y <= p.y && p.y <= bmax.y)) return 1;
                }
            }
        }
    }
    return 0;
}
int orc_box_intersect(const double *bmin, const double *bmax, const double *o, const double *d) {
    return box_intersect(v_ld(bmin), v_ld(bmax), v_ld(o), v_ld(d));
}

/* ============================ shared/state/camera.go ======================== */
/* camera.go:35-44; GlobalUp = (0,1,0) environment.go:22 */
int orc_new_camera(const double *dir, double *basis9) {
    vec gup = { 0, 1, 0 };
    vec d = v_ld(dir);
    if (v_zero(v_cross(d, gup))) return -1;
    vec forward = v_norm(d);
    vec left = v_norm(v_cross(d, gup));
    vec up = v_cross(left, forward);
    v_st(basis9, forward); v_st(basis9 + 3, left); v_st(basis9 + 6, up);
    return 0;
}

/* tracer.go:15-22 with tan(fov/2) supplied (frame constant), then :86 */
static inline vec primary_dir(int i, int j, int W, int H, vec pos, vec fwd, vec left, vec up, double projHalfWidth) {
    int halfWidth = W / 2, halfHeight = H / 2;                                                 /* :16 */
    double projHalfHeight = projHalfWidth * (double)H / (double)W;                             /* :18 */
    vec iOff = v_scale(left, projHalfWidth * ((double)(halfWidth - i) - 0.5) / (double)halfWidth);   /* :19 */
    vec jOff = v_scale(up, projHalfHeight * ((double)(halfHeight - j) - 0.5) / (double)halfHeight); /* :20 */
    vec pt = v_add(v_add(v_add(pos, fwd), iOff), jOff);                                        /* :21 */
    return v_norm(v_sub(pt, pos));                                                             /* :86 */
}
void orc_primary_ray(int i, int j, int W, int H, const double *c, double *dir3) {
    v_st(dir3, primary_dir(i, j, W, H, v_ld(c), v_ld(c + 3), v_ld(c + 6), v_ld(c + 9), orc_go_tan(c[12] / 2.0)));
}

/* ============================ scene ========================================= */
typedef struct { vec bmin, bmax; int32_t left, right, first, count; } cnode;   /* culling BVH node */

/* Guttman R-tree restated from the published algorithm (rtreego is absent; unpinned). */
typedef struct rnode {
    int leaf, n;
    vec emin[6], emax[6];        /* entry rects: emin = p, emax = p + lengths (rtreego Rect{p,q}) */
    struct rnode *child[6];
    int32_t face[6];
    struct rnode *parent;
} rnode;

typedef struct {
    uint32_t nv, nn, nf, nm;
    vec *verts, *normals;
    uint32_t *faces;
    double *materials;
    /* culling BVH */
    cnode *nodes; int32_t n_nodes; int32_t *order;
    /* reference-shaped R-tree */
    rnode *rroot;
} omesh;

struct orc_scene {
    uint32_t n_meshes, n_obj, n_lights;
    omesh *meshes;
    vec *obj_pos; uint32_t *obj_mesh;
    vec *obj_bmin, *obj_bmax;     /* object.go:31-59 */
    vec cam_pos, cam_fwd, cam_left, cam_up; double fov, tan_half;
    vec *light_pos, *light_col;
};

/* ---- culling BVH (oracle-internal accelerator; may only cull, never reorder) ---- */
static void face_bounds(const omesh *m, uint32_t f, vec *lo, vec *hi) {
    const uint32_t *fc = m->faces + 7 * (size_t)f;
    vec a = m->verts[fc[0]], b = m->verts[fc[1]], c = m->verts[fc[2]];
    lo->x = fmin(a.x, fmin(b.x, c.x)); hi->x = fmax(a.x, fmax(b.x, c.x));
    lo->y = fmin(a.y, fmin(b.y, c.y)); hi->y = fmax(a.y, fmax(b.y, c.y));
    lo->z = fmin(a.z, fmin(b.z, c.z)); hi->z = fmax(a.z, fmax(b.z, c.z));
}

/* Median partition of order[first .. first+count) by (key, index): after the call every entry of the lower half compares
 * <= every entry of the upper half.  Quickselect (median of three, Hoare partition) -- O(n) per level instead of a full sort;
 * the tree may only CULL (padded boxes, no tmax), so any split is a valid one: only build time depends on it (a 10 M-face soup
 * took 116 s with qsort at every level). */
static inline int key_less(const double *key, int32_t a, int32_t b) { return key[a] < key[b] || (key[a] == key[b] && a < b); }
static void median_partition(int32_t *v, int32_t n, int32_t k, const double *key) {
    int32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        /* median of three -> pivot value */
        int32_t a = v[lo], b = v[mid], c = v[hi], p;
        if (key_less(key, a, b)) p = key_less(key, b, c) ? b : (key_less(key, a, c) ? c : a);
        else p = key_less(key, a, c) ? a : (key_less(key, b, c) ? c : b);
        int32_t i = lo, j = hi;
        while (i <= j) {
            while (key_less(key, v[i], p)) i++;
            while (key_less(key, p, v[j])) j--;
            if (i <= j) { int32_t t = v[i]; v[i] = v[j]; v[j] = t; i++; j--; }
        }
        if (k <= j) hi = j;
        else if (k >= i) lo = i;
        else return;
    }
}

/* Subtrees of more than CBVH_PAR faces are built by their own thread (the node array is preallocated, ids come from an atomic
 * counter): the tree only culls, so the order in which ids are handed out is irrelevant to every result. */
#define CBVH_PAR 400000
typedef struct { omesh *m; const vec *flo, *fhi; double **cent; int32_t first, count; double pad; int32_t id; } cbvh_job;
static int32_t cbvh_build(omesh *m, const vec *flo, const vec *fhi, double *cent[3], int32_t first, int32_t count, double pad);
static void *cbvh_thread(void *arg) {
    cbvh_job *j = (cbvh_job *)arg;
    j->id = cbvh_build(j->m, j->flo, j->fhi, j->cent, j->first, j->count, j->pad);
    return NULL;
}

static int32_t cbvh_build(omesh *m, const vec *flo, const vec *fhi, double *cent[3], int32_t first, int32_t count,
                          double pad) {
    int32_t id = __atomic_fetch_add(&m->n_nodes, 1, __ATOMIC_RELAXED);
    cnode *n = &m->nodes[id];
    vec lo = { DBL_MAX, DBL_MAX, DBL_MAX }, hi = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    vec clo = lo, chi = hi;
    for (int32_t k = first; k < first + count; k++) {
        int32_t f = m->order[k];
        lo.x = fmin(lo.x, flo[f].x); lo.y = fmin(lo.y, flo[f].y); lo.z = fmin(lo.z, flo[f].z);
        hi.x = fmax(hi.x, fhi[f].x); hi.y = fmax(hi.y, fhi[f].y); hi.z = fmax(hi.z, fhi[f].z);
        clo.x = fmin(clo.x, cent[0][f]); clo.y = fmin(clo.y, cent[1][f]); clo.z = fmin(clo.z, cent[2][f]);
        chi.x = fmax(chi.x, cent[0][f]); chi.y = fmax(chi.y, cent[1][f]); chi.z = fmax(chi.z, cent[2][f]);
    }
    n->bmin.x = lo.x - pad; n->bmin.y = lo.y - pad; n->bmin.z = lo.z - pad;
    n->bmax.x = hi.x + pad; n->bmax.y = hi.y + pad; n->bmax.z = hi.z + pad;
    n->first = first; n->count = count; n->left = n->right = -1;
    if (count <= 4) return id;
    double ex = chi.x - clo.x, ey = chi.y - clo.y, ez = chi.z - clo.z;
    int ax = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
    int32_t half = count / 2;
    median_partition(m->order + first, count, half, cent[ax]);
    int32_t l, r;
    pthread_t th;
    cbvh_job job = { m, flo, fhi, cent, first, half, pad, -1 };
    if (count > CBVH_PAR && pthread_create(&th, NULL, cbvh_thread, &job) == 0) {
        r = cbvh_build(m, flo, fhi, cent, first + half, count - half, pad);
        pthread_join(th, NULL);
        l = job.id;
    } else {
        l = cbvh_build(m, flo, fhi, cent, first, half, pad);
        r = cbvh_build(m, flo, fhi, cent, first + half, count - half, pad);
    }
    n->left = l; n->right = r; n->count = 0;
    return id;
}

/* conservative fp64 slab test against a padded box, no tmax (order-independent) */
static inline int slab_hit(vec bmin, vec bmax, vec o, vec d) {
    double t0 = 0.0, t1 = DBL_MAX;
    const double oo[3] = { o.x, o.y, o.z }, dd[3] = { d.x, d.y, d.z };
    const double lo[3] = { bmin.x, bmin.y, bmin.z }, hi[3] = { bmax.x, bmax.y, bmax.z };
    for (int a = 0; a < 3; a++) {
        if (dd[a] == 0.0) {
            if (oo[a] < lo[a] || oo[a] > hi[a]) return 0;
        } else {
            double ta = (lo[a] - oo[a]) / dd[a], tb = (hi[a] - oo[a]) / dd[a];
            if (ta > tb) { double s = ta; ta = tb; tb = s; }
            /* widen by a relative 1e-12 so rounding in the divisions can never exclude a true hit */
            ta -= fabs(ta) * 1e-12; tb += fabs(tb) * 1e-12;
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
            if (t0 > t1) return 0;
        }
    }
    return 1;
}

/* ---- R-tree (Guttman 1984, quadratic split; parameters NewTree(3,2,5) mesh.go:139) ---- */
#define RT_MIN 2
#define RT_MAX 5
static double rect_size(vec lo, vec hi) { return (hi.x - lo.x) * (hi.y - lo.y) * (hi.z - lo.z); }
static void rect_union(vec alo, vec ahi, vec blo, vec bhi, vec *lo, vec *hi) {
    lo->x = fmin(alo.x, blo.x); lo->y = fmin(alo.y, blo.y); lo->z = fmin(alo.z, blo.z);
    hi->x = fmax(ahi.x, bhi.x); hi->y = fmax(ahi.y, bhi.y); hi->z = fmax(ahi.z, bhi.z);
}
static void rnode_bb(const rnode *n, vec *lo, vec *hi) {
    *lo = n->emin[0]; *hi = n->emax[0];
    for (int k = 1; k < n->n; k++) rect_union(*lo, *hi, n->emin[k], n->emax[k], lo, hi);
}
static rnode *rnode_new(int leaf) { rnode *n = (rnode *)calloc(1, sizeof(rnode)); n->leaf = leaf; return n; }

static rnode *rt_choose_leaf(rnode *n, vec lo, vec hi, int level, int cur_level) {
    if (n->leaf || cur_level == level) return n;
    double diff = DBL_MAX; int chosen = 0;
    for (int k = 0; k < n->n; k++) {
        vec ulo, uhi;
        rect_union(n->emin[k], n->emax[k], lo, hi, &ulo, &uhi);
        double sz = rect_size(n->emin[k], n->emax[k]);
        double dd = rect_size(ulo, uhi) - sz;
        if (dd < diff || (dd == diff && sz < rect_size(n->emin[chosen], n->emax[chosen]))) { diff = dd; chosen = k; }
    }
    return rt_choose_leaf(n->child[chosen], lo, hi, level, cur_level - 1);
}

typedef struct { vec lo, hi; rnode *child; int32_t face; } rentry;
static void rnode_push(rnode *n, rentry e) {
    n->emin[n->n] = e.lo; n->emax[n->n] = e.hi; n->child[n->n] = e.child; n->face[n->n] = e.face;
    if (e.child) e.child->parent = n;
    n->n++;
}

/* quadratic split; left re-uses n */
static rnode *rt_split(rnode *n) {
    rentry es[RT_MAX + 1];
    int cnt = n->n;
    for (int k = 0; k < cnt; k++) { es[k].lo = n->emin[k]; es[k].hi = n->emax[k]; es[k].child = n->child[k]; es[k].face = n->face[k]; }
    /* pickSeeds: pair wasting the most volume */
    int l = 0, r = 1; double maxw = -DBL_MAX;
    for (int a = 0; a < cnt; a++) for (int b = a + 1; b < cnt; b++) {
        vec ulo, uhi; rect_union(es[a].lo, es[a].hi, es[b].lo, es[b].hi, &ulo, &uhi);
        double w = rect_size(ulo, uhi) - rect_size(es[a].lo, es[a].hi) - rect_size(es[b].lo, es[b].hi);
        if (w > maxw) { maxw = w; l = a; r = b; }
    }
    rnode *right = rnode_new(n->leaf);
    right->parent = n->parent;
    n->n = 0;
    rnode_push(n, es[l]); rnode_push(right, es[r]);
    int used[RT_MAX + 1] = { 0 }; used[l] = used[r] = 1;
    int remaining = cnt - 2;
    while (remaining > 0) {
        vec llo, lhi, rlo, rhi; rnode_bb(n, &llo, &lhi); rnode_bb(right, &rlo, &rhi);
        /* pickNext: max |d1-d2| */
        int next = -1; double best = -1.0;
        for (int k = 0; k < cnt; k++) if (!used[k]) {
            vec ulo, uhi;
            rect_union(llo, lhi, es[k].lo, es[k].hi, &ulo, &uhi); double d1 = rect_size(ulo, uhi) - rect_size(llo, lhi);
            rect_union(rlo, rhi, es[k].lo, es[k].hi, &ulo, &uhi); double d2 = rect_size(ulo, uhi) - rect_size(rlo, rhi);
            double dd = fabs(d1 - d2);
            if (dd > best) { best = dd; next = k; }
        }
        rentry e = es[next];
        if (remaining + n->n <= RT_MIN) rnode_push(n, e);
        else if (remaining + right->n <= RT_MIN) rnode_push(right, e);
        else {
            vec ulo, uhi;
            rect_union(llo, lhi, e.lo, e.hi, &ulo, &uhi); double gl = rect_size(ulo, uhi) - rect_size(llo, lhi);
            rect_union(rlo, rhi, e.lo, e.hi, &ulo, &uhi); double gr = rect_size(ulo, uhi) - rect_size(rlo, rhi);
            if (gl < gr) rnode_push(n, e);
            else if (gr < gl) rnode_push(right, e);
            else {
                double sl = rect_size(llo, lhi), sr = rect_size(rlo, rhi);
                if (sl < sr) rnode_push(n, e);
                else if (sr < sl) rnode_push(right, e);
                else if (n->n <= right->n) rnode_push(n, e);
                else rnode_push(right, e);
            }
        }
        used[next] = 1; remaining--;
    }
    return right;
}

static void rt_insert(rnode **root, int *height, vec lo, vec hi, int32_t face) {
    rnode *leaf = rt_choose_leaf(*root, lo, hi, 1, *height);
    rentry e = { lo, hi, NULL, face };
    rnode_push(leaf, e);
    rnode *n = leaf, *split = NULL;
    if (n->n > RT_MAX) split = rt_split(n);
    /* adjustTree */
    while (n->parent) {
        rnode *p = n->parent;
        for (int k = 0; k < p->n; k++) if (p->child[k] == n) { rnode_bb(n, &p->emin[k], &p->emax[k]); break; }
        rnode *psplit = NULL;
        if (split) {
            rentry se; rnode_bb(split, &se.lo, &se.hi); se.child = split; se.face = -1;
            rnode_push(p, se);
            if (p->n > RT_MAX) psplit = rt_split(p);
        }
        n = p; split = psplit;
    }
    if (split) {
        rnode *nr = rnode_new(0);
        rentry a, b;
        rnode_bb(n, &a.lo, &a.hi); a.child = n; a.face = -1;
        rnode_bb(split, &b.lo, &b.hi); b.child = split; b.face = -1;
        rnode_push(nr, a); rnode_push(nr, b);
        *root = nr; (*height)++;
    }
}
static void rt_free(rnode *n) {
    if (!n) return;
    if (!n->leaf) for (int k = 0; k < n->n; k++) rt_free(n->child[k]);
    free(n);
}

static void mesh_build_rtree(omesh *m) {
    rnode *root = rnode_new(1); int height = 1;
    for (uint32_t f = 0; f < m->nf; f++) {
        vec lo, hi; face_bounds(m, f, &lo, &hi);
        /* mesh.go:44 lengths = max(max-min, boundEpsilon); rtreego Rect q = p + len; box.go:23-24 max = p + (q-p) */
        vec len = { go_max(hi.x - lo.x, 1e-4), go_max(hi.y - lo.y, 1e-4), go_max(hi.z - lo.z, 1e-4) };
        vec q = { lo.x + len.x, lo.y + len.y, lo.z + len.z };
        vec mx = { lo.x + (q.x - lo.x), lo.y + (q.y - lo.y), lo.z + (q.z - lo.z) };
        rt_insert(&root, &height, lo, mx, (int32_t)f);
    }
    m->rroot = root;
}

orc_scene *orc_scene_create(const orc_mesh_desc *meshes, uint32_t n_meshes, const double *obj_pos,
                            const uint32_t *obj_mesh, uint32_t n_obj) {
    orc_scene *s = (orc_scene *)calloc(1, sizeof(*s));
    s->n_meshes = n_meshes; s->n_obj = n_obj;
    s->meshes = (omesh *)calloc(n_meshes ? n_meshes : 1, sizeof(omesh));
    for (uint32_t k = 0; k < n_meshes; k++) {
        omesh *m = &s->meshes[k]; const orc_mesh_desc *d = &meshes[k];
        m->nv = d->n_verts; m->nn = d->n_normals; m->nf = d->n_faces; m->nm = d->n_materials;
        m->verts = (vec *)malloc(sizeof(vec) * (m->nv ? m->nv : 1));
        m->normals = (vec *)malloc(sizeof(vec) * (m->nn ? m->nn : 1));
        m->faces = (uint32_t *)malloc(sizeof(uint32_t) * 7 * (m->nf ? m->nf : 1));
        m->materials = (double *)malloc(sizeof(double) * 10 * (m->nm ? m->nm : 1));
        for (uint32_t i = 0; i < m->nv; i++) m->verts[i] = v_ld(d->verts + 3 * (size_t)i);
        for (uint32_t i = 0; i < m->nn; i++) m->normals[i] = v_ld(d->normals + 3 * (size_t)i);
        memcpy(m->faces, d->faces, sizeof(uint32_t) * 7 * (size_t)m->nf);
        memcpy(m->materials, d->materials, sizeof(double) * 10 * (size_t)m->nm);
        /* culling BVH */
        if (m->nf) {
            vec *flo = (vec *)malloc(sizeof(vec) * m->nf), *fhi = (vec *)malloc(sizeof(vec) * m->nf);
            double *cent[3];
            for (int a = 0; a < 3; a++) cent[a] = (double *)malloc(sizeof(double) * m->nf);
            vec lo = { DBL_MAX, DBL_MAX, DBL_MAX }, hi = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
            for (uint32_t f = 0; f < m->nf; f++) {
                face_bounds(m, f, &flo[f], &fhi[f]);
                cent[0][f] = 0.5 * (flo[f].x + fhi[f].x); cent[1][f] = 0.5 * (flo[f].y + fhi[f].y); cent[2][f] = 0.5 * (flo[f].z + fhi[f].z);
                lo.x = fmin(lo.x, flo[f].x); lo.y = fmin(lo.y, flo[f].y); lo.z = fmin(lo.z, flo[f].z);
                hi.x = fmax(hi.x, fhi[f].x); hi.y = fmax(hi.y, fhi[f].y); hi.z = fmax(hi.z, fhi[f].z);
            }
            double ext = fmax(hi.x - lo.x, fmax(hi.y - lo.y, hi.z - lo.z));
            double mag = fmax(fmax(fabs(lo.x), fabs(hi.x)), fmax(fmax(fabs(lo.y), fabs(hi.y)), fmax(fabs(lo.z), fabs(hi.z))));
            double pad = 1e-7 * fmax(ext, mag) + 1e-30;
            m->order = (int32_t *)malloc(sizeof(int32_t) * m->nf);
            for (uint32_t f = 0; f < m->nf; f++) m->order[f] = (int32_t)f;
            m->nodes = (cnode *)malloc(sizeof(cnode) * (2 * (size_t)m->nf + 1));
            m->n_nodes = 0;
            cbvh_build(m, flo, fhi, cent, 0, (int32_t)m->nf, pad);
            free(flo); free(fhi);
            for (int a = 0; a < 3; a++) free(cent[a]);
        }
        m->rroot = NULL; /* built lazily by orc_build_rtrees */
    }
    s->obj_pos = (vec *)malloc(sizeof(vec) * (n_obj ? n_obj : 1));
    s->obj_mesh = (uint32_t *)malloc(sizeof(uint32_t) * (n_obj ? n_obj : 1));
    s->obj_bmin = (vec *)malloc(sizeof(vec) * (n_obj ? n_obj : 1));
    s->obj_bmax = (vec *)malloc(sizeof(vec) * (n_obj ? n_obj : 1));
    for (uint32_t k = 0; k < n_obj; k++) { s->obj_pos[k] = v_ld(obj_pos + 3 * k); s->obj_mesh[k] = obj_mesh[k]; }
    s->tan_half = NAN;
    return s;
}

void orc_scene_destroy(orc_scene *s) {
    if (!s) return;
    for (uint32_t k = 0; k < s->n_meshes; k++) {
        omesh *m = &s->meshes[k];
        free(m->verts); free(m->normals); free(m->faces); free(m->materials); free(m->nodes); free(m->order);
        rt_free(m->rroot);
    }
    free(s->meshes); free(s->obj_pos); free(s->obj_mesh); free(s->obj_bmin); free(s->obj_bmax);
    free(s->light_pos); free(s->light_col);
    free(s);
}

/* object.go:31-59 : object bounds (includes the object's own Pos; extents clamped to boundEpsilon);
 * rtreego Rect{p, q=p+len}; box.go:21-26 max = p + (q-p). */
static void object_bounds(orc_scene *s, uint32_t k) {
    vec P = s->obj_pos[k];
    const omesh *m = &s->meshes[s->obj_mesh[k]];
    double xMin = P.x, xMax = P.x, yMin = P.y, yMax = P.y, zMin = P.z, zMax = P.z;
    for (uint32_t i = 0; i < m->nv; i++) {
        vec v = m->verts[i];
        xMin = go_min(xMin, P.x + v.x); xMax = go_max(xMax, P.x + v.x);
        yMin = go_min(yMin, P.y + v.y); yMax = go_max(yMax, P.y + v.y);
        zMin = go_min(zMin, P.z + v.z); zMax = go_max(zMax, P.z + v.z);
    }
    vec len = { go_max(xMax - xMin, 1e-4), go_max(yMax - yMin, 1e-4), go_max(zMax - zMin, 1e-4) };
    vec p = { xMin, yMin, zMin };
    vec q = { p.x + len.x, p.y + len.y, p.z + len.z };
    s->obj_bmin[k] = p;
    vec mx = { p.x + (q.x - p.x), p.y + (q.y - p.y), p.z + (q.z - p.z) };
    s->obj_bmax[k] = mx;
}

void orc_frame_set(orc_scene *s, const double *c, double tan_half_fov, const double *lights, uint32_t n_lights,
                   const double *obj_pos) {
    s->cam_pos = v_ld(c); s->cam_fwd = v_ld(c + 3); s->cam_left = v_ld(c + 6); s->cam_up = v_ld(c + 9);
    s->fov = c[12];
    s->tan_half = isnan(tan_half_fov) ? orc_go_tan(s->fov / 2.0) : tan_half_fov;
    free(s->light_pos); free(s->light_col);
    s->n_lights = n_lights;
    s->light_pos = (vec *)malloc(sizeof(vec) * (n_lights ? n_lights : 1));
    s->light_col = (vec *)malloc(sizeof(vec) * (n_lights ? n_lights : 1));
    for (uint32_t k = 0; k < n_lights; k++) { s->light_pos[k] = v_ld(lights + 6 * k); s->light_col[k] = v_ld(lights + 6 * k + 3); }
    if (obj_pos) for (uint32_t k = 0; k < s->n_obj; k++) s->obj_pos[k] = v_ld(obj_pos + 3 * k);
    for (uint32_t k = 0; k < s->n_obj; k++) object_bounds(s, k);
}

void orc_build_rtrees(orc_scene *s) {
    for (uint32_t k = 0; k < s->n_meshes; k++) if (!s->meshes[k].rroot && s->meshes[k].nf) mesh_build_rtree(&s->meshes[k]);
}

/* ---- object.go:63-110 ------------------------------------------------------ */
typedef struct {
    int has; double dist; vec ip; double r1, r2, r3; int32_t face;
} nearest_t;

static inline void consider_face(const omesh *m, int32_t f, vec o, vec d, nearest_t *nr, orc_stats *st) {
    const uint32_t *fc = m->faces + 7 * (size_t)f;
    vec ip; double r1, r2, r3, t;
    st->tri_tests++;
    if (tri_intersection(m->verts[fc[0]], m->verts[fc[1]], m->verts[fc[2]], o, d, &ip, &r1, &r2, &r3, &t)) {   /* :88 */
        double dist = v_len(v_sub(o, ip));                                                                       /* :97 */
        /* :98 strict '<' over candidates in enumeration order == lexicographic (dist, face index) minimum */
        if (nr->has && dist == nr->dist && f != nr->face) st->nearest_ties++;
        if (!nr->has || dist < nr->dist || (dist == nr->dist && f < nr->face)) {
            nr->has = 1; nr->dist = dist; nr->ip = ip; nr->r1 = r1; nr->r2 = r2; nr->r3 = r3; nr->face = f;
        }
    }
}

static void rt_search(const omesh *m, const rnode *n, vec o, vec d, nearest_t *nr, orc_stats *st) {
    for (int k = 0; k < n->n; k++) {
        st->box_tests++;
        if (!box_intersect(n->emin[k], n->emax[k], o, d)) continue;
        if (n->leaf) consider_face(m, n->face[k], o, d, nr, st);
        else rt_search(m, n->child[k], o, d, nr, st);
    }
}

/* Diagnostic (tests / tools only): follows the R-tree path from the root to `face` of mesh `mesh` and evaluates box.go's
 * predicate on every entry rectangle of that path for the OBJECT-SPACE ray (o, d).  Returns the number of rectangles on the
 * path (0: face not found / no tree); *first_rejected = index of the first rectangle the predicate rejects (-1: none), and
 * that rectangle in bmin / bmax.  Used to show WHY the reference-shaped walk can miss a real hit. */
static int rt_path(const rnode *n, int32_t face, const rnode **nodes, int *slots, int depth) {
    for (int k = 0; k < n->n; k++) {
        nodes[depth] = n; slots[depth] = k;
        if (n->leaf) { if (n->face[k] == face) return depth + 1; }
        else { int r = rt_path(n->child[k], face, nodes, slots, depth + 1); if (r) return r; }
    }
    return 0;
}
int orc_rtree_path_check(const orc_scene *s, uint32_t mesh, int32_t face, const double *o, const double *d, int *first_rejected,
                         double *bmin, double *bmax) {
    *first_rejected = -1;
    if (mesh >= s->n_meshes || !s->meshes[mesh].rroot) return 0;
    const rnode *nodes[64]; int slots[64];
    int len = rt_path(s->meshes[mesh].rroot, face, nodes, slots, 0);
    for (int k = 0; k < len; k++) {
        vec lo = nodes[k]->emin[slots[k]], hi = nodes[k]->emax[slots[k]];
        if (!box_intersect(lo, hi, v_ld(o), v_ld(d))) { *first_rejected = k; v_st(bmin, lo); v_st(bmax, hi); break; }
    }
    return len;
}

/* Returns hit; ip_world, normal, face index, material index */
static int object_intersection(const orc_scene *s, int mode, uint32_t k, vec o, vec d, vec *ipw, vec *normal,
                               int32_t *face, int32_t *mat, orc_stats *st) {
    const omesh *m = &s->meshes[s->obj_mesh[k]];
    vec P = s->obj_pos[k];
    o = v_sub(o, P);                                            /* :71 */
    nearest_t nr; nr.has = 0; nr.dist = 0; nr.face = -1;
    if (m->nf == 0) return 0;
    if (mode == ORC_BRUTE) {
        for (int32_t f = 0; f < (int32_t)m->nf; f++) consider_face(m, f, o, d, &nr, st);
    } else if (mode == ORC_RTREE && m->rroot) {
        rt_search(m, m->rroot, o, d, &nr, st);
    } else {
        int32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const cnode *n = &m->nodes[stack[--sp]];
            st->box_tests++;
            if (!slab_hit(n->bmin, n->bmax, o, d)) continue;
            if (n->left < 0) { for (int32_t q = n->first; q < n->first + n->count; q++) consider_face(m, m->order[q], o, d, &nr, st); }
            else { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
    }
    if (!nr.has) return 0;
    const uint32_t *fc = m->faces + 7 * (size_t)nr.face;
    vec nrm;
    if (m->nn > 0) {
        /* triangle.go:29-31 InterpNormal */
        vec N1 = m->normals[fc[3]], N2 = m->normals[fc[4]], N3 = m->normals[fc[5]];
        nrm = v_norm(v_add(v_add(v_scale(N1, nr.r1), v_scale(N2, nr.r2)), v_scale(N3, nr.r3)));
    } else {
        /* triangle.go:24-26 Normal */
        vec P1 = m->verts[fc[0]], P2 = m->verts[fc[1]], P3 = m->verts[fc[2]];
        nrm = v_norm(v_cross(v_sub(P2, P1), v_sub(P3, P1)));
    }
    *ipw = v_add(nr.ip, P);                                     /* :109 */
    *normal = nrm; *face = nr.face; *mat = (int32_t)fc[6];
    return 1;
}

/* tracer.go:27-50.  Objects are visited in scene order (object index); the reference visits them in
 * object-R-tree enumeration order, which matters only for exact ties of the camera distance. */
static int trace_ray(const orc_scene *s, int mode, vec o, vec d, vec *ip, vec *normal, int32_t *face, int32_t *obj,
                     int32_t *mat, orc_stats *st) {
    int exists = 0; double nd = 0;
    for (uint32_t k = 0; k < s->n_obj; k++) {
        /* tracer.go:32 : object-level candidate filter (box.go predicate on object.go:31-59 bounds).
         * Applied only in the reference-shaped mode; BRUTE/CULLED test every object (superset). */
        if (mode == ORC_RTREE) { st->box_tests++; if (!box_intersect(s->obj_bmin[k], s->obj_bmax[k], o, d)) continue; }
        vec p, n; int32_t f, m;
        if (object_intersection(s, mode, k, o, d, &p, &n, &f, &m, st)) {
            double dist = v_len(v_sub(p, s->cam_pos));           /* :38 (camera position, even for shadow rays) */
            if (!exists || dist < nd) {                          /* :39 */
                exists = 1; nd = dist; *ip = p; *normal = n; *face = f; *obj = (int32_t)k; *mat = m;
            }
        }
    }
    return exists;
}

int orc_trace(const orc_scene *s, int mode, const double *o, const double *d, double *ip, double *normal,
              int32_t *face, int32_t *obj, int32_t *mat) {
    orc_stats st; memset(&st, 0, sizeof st);
    vec p, n;
    int hit = trace_ray(s, mode, v_ld(o), v_ld(d), &p, &n, face, obj, mat, &st);
    if (hit) { v_st(ip, p); v_st(normal, n); }
    return hit;
}

/* colour.go:38-50 */
static inline double c_add(double a, double b) { return go_min(a + b, 1.0); }
static inline double c_scale(double a, double s) { return go_max(0.0, go_min(s * a, 1.0)); }

/* tracer.go:53-77 */
static void phong(const orc_scene *s, int mode, vec P, vec N, const double *mat10, double *col, uint32_t *lit_mask,
                  orc_stats *st) {
    double c[3] = { mat10[0], mat10[1], mat10[2] };                       /* :56 colour := material.Ka */
    const double *Kd = mat10 + 3, *Ks = mat10 + 6; double Ns = mat10[9];
    uint32_t lm = 0;
    for (uint32_t l = 0; l < s->n_lights; l++) {
        vec Lpos = s->light_pos[l], Lcol = s->light_col[l];
        vec lightDir = v_norm(v_sub(Lpos, P));                            /* :61 */
        vec sp, sn; int32_t sf, so, sm;
        st->shadow_rays++;
        int shaded = trace_ray(s, mode, v_add(P, v_scale(lightDir, 0.0001)), lightDir, &sp, &sn, &sf, &so, &sm, st); /* :64 */
        if (!shaded || v_len(v_sub(Lpos, P)) < v_len(v_sub(sp, P))) {
            lm |= 1u << (l & 31);
            vec reflectDir = v_sub(v_scale(N, 2 * v_dot(lightDir, N)), lightDir);       /* :65 */
            vec camDir = v_norm(v_sub(s->cam_pos, P));                                   /* :66 */
            double diff = go_max(v_dot(lightDir, N), 0.0);                               /* :69 */
            double spec = orc_go_pow(go_max(v_dot(reflectDir, camDir), 0.0), Ns);        /* :72 */
            const double lc[3] = { Lcol.x, Lcol.y, Lcol.z };
            for (int k = 0; k < 3; k++) c[k] = c_add(c[k], c_scale(Kd[k], diff) * lc[k]);  /* :69 */
            for (int k = 0; k < 3; k++) c[k] = c_add(c[k], c_scale(Ks[k], spec) * lc[k]);  /* :72 */
        }
    }
    col[0] = c[0]; col[1] = c[1]; col[2] = c[2];
    *lit_mask = lm;
}

static int trace_pixel(const orc_scene *s, int mode, int i, int j, int W, int H, double *col, int32_t *face,
                       int32_t *obj, uint32_t *lit, orc_stats *st) {
    vec dir = primary_dir(i, j, W, H, s->cam_pos, s->cam_fwd, s->cam_left, s->cam_up, s->tan_half);
    vec ip, n; int32_t f = -1, o = -1, m = -1;
    st->primary_rays++;
    if (trace_ray(s, mode, s->cam_pos, dir, &ip, &n, &f, &o, &m, st)) {          /* tracer.go:86 */
        st->hit_pixels++;
        const omesh *mm = &s->meshes[s->obj_mesh[o]];
        phong(s, mode, ip, n, mm->materials + 10 * (size_t)m, col, lit, st);      /* :87 */
        *face = f; *obj = o;
        return 1;
    }
    col[0] = col[1] = col[2] = 0; *face = -1; *obj = -1; *lit = 0;               /* :89 */
    return 0;
}

int orc_trace_pixel(const orc_scene *s, int mode, int i, int j, int W, int H, double *colour, int32_t *face,
                    int32_t *obj, uint32_t *lit_mask) {
    orc_stats st; memset(&st, 0, sizeof st);
    return trace_pixel(s, mode, i, j, W, H, colour, face, obj, lit_mask, &st);
}

static inline uint8_t quantise(double c, int quant) {
    if (quant == ORC_Q_RGBA16_SHR8) return (uint8_t)(((uint32_t)(65535.0 * c)) >> 8);   /* colour.go:54-56 + go-sdl2 Set */
    return (uint8_t)(255 * c);                                                          /* colour.go:59-61 */
}

typedef struct {
    const orc_scene *s; int mode, W, H, x, y, w, h, layout, quant, step, phase;
    uint8_t *out_rgb; double *out_colour; int32_t *out_face;
    int *next_col; pthread_mutex_t *mu; orc_stats *total;
} region_job;

static void *region_worker(void *arg) {
    region_job *jb = (region_job *)arg;
    orc_stats st; memset(&st, 0, sizeof st);
    const int w = jb->w, h = jb->h;
    for (;;) {
        int i = __atomic_fetch_add(jb->next_col, 1, __ATOMIC_RELAXED);    /* distributed/main.go:67 */
        if (i >= w) break;
        for (int j = 0; j < h; j++) {                                     /* :68 */
            long lin = (long)i * h + j;
            if (jb->step > 1 && (lin % jb->step) != jb->phase) continue;
            double col[3]; int32_t f, o; uint32_t lit;
            trace_pixel(jb->s, jb->mode, jb->x + i, jb->y + j, jb->W, jb->H, col, &f, &o, &lit, &st);   /* :78 */
            size_t idx = (jb->layout == ORC_L_COLMAJOR) ? (size_t)i * h + j : (size_t)j * w + i;      /* :82 */
            if (jb->out_rgb) for (int k = 0; k < 3; k++) jb->out_rgb[3 * idx + k] = quantise(col[k], jb->quant);
            if (jb->out_colour) for (int k = 0; k < 3; k++) jb->out_colour[3 * idx + k] = col[k];
            if (jb->out_face) jb->out_face[idx] = f;
        }
    }
    pthread_mutex_lock(jb->mu);
    orc_stats *t = jb->total;
    t->primary_rays += st.primary_rays; t->shadow_rays += st.shadow_rays; t->hit_pixels += st.hit_pixels;
    t->tri_tests += st.tri_tests; t->box_tests += st.box_tests; t->nearest_ties += st.nearest_ties;
    pthread_mutex_unlock(jb->mu);
    return NULL;
}

int orc_render_region(const orc_scene *s, int mode, int W, int H, int x, int y, int w, int h, int layout, int quant,
                      int n_threads, int step, int phase, uint8_t *out_rgb, double *out_colour, int32_t *out_face,
                      orc_stats *stats) {
    orc_stats total; memset(&total, 0, sizeof total);
    if (step < 1) step = 1;
    if (mode == ORC_RTREE) orc_build_rtrees((orc_scene *)s);
    if (n_threads <= 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); n_threads = nc > 0 ? (int)nc : 1; }
    if (n_threads > 256) n_threads = 256;
    if (n_threads > w && w > 0) n_threads = w;
    int next_col = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    region_job jb = { s, mode, W, H, x, y, w, h, layout, quant, step, phase, out_rgb, out_colour, out_face,
                      &next_col, &mu, &total };
    if (n_threads <= 1) {
        region_worker(&jb);
    } else {
        pthread_t th[256];
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, region_worker, &jb);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    if (stats) *stats = total;
    return 0;
}

/* ---- comms.TraceResults wire bytes (see rt_oracle.h) -------------------------------------------- */
static uint64_t put_varint(uint8_t *out, uint64_t pos, uint32_t v) {
    while (v >= 0x80u) {
        if (out) out[pos] = (uint8_t)((v & 0x7fu) | 0x80u);
        pos++;
        v >>= 7;
    }
    if (out) out[pos] = (uint8_t)v;
    return pos + 1;
}

static uint32_t varint_len(uint32_t v) {
    uint32_t n = 1;
    while (v >= 0x80u) { n++; v >>= 7; }
    return n;
}

uint64_t orc_encode_trace_results(const uint8_t *rgb, uint64_t n, uint8_t *out) {
    uint64_t pos = 0;
    for (uint64_t p = 0; p < n; p++) {
        uint32_t len = 0;
        for (int c = 0; c < 3; c++)
            if (rgb[3 * p + c]) len += 1 + varint_len(rgb[3 * p + c]);     /* proto3: default (0) scalars are not emitted */
        if (out) out[pos] = 0x0A;                                          /* field 1, wire type 2 */
        pos++;
        pos = put_varint(out, pos, len);
        for (int c = 0; c < 3; c++) {
            uint32_t v = rgb[3 * p + c];
            if (!v) continue;
            if (out) out[pos] = (uint8_t)((c + 1) << 3);                   /* field c+1, wire type 0 */
            pos++;
            pos = put_varint(out, pos, v);
        }
    }
    return pos;
}
