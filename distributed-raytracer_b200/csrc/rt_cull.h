// rt_cull.h -- primary-ray culling rectangle (host side, per frame; a few hundred flops).
//
// ~90 % of the primary rays of the BASELINE views miss every object.  The reference finds that out per ray
// (tracer.go:32: object R-tree + box.go:29-69); here the padded root box of every object is projected through the camera
// basis once per frame and the primary kernels are launched only over the 8x4 pixel blocks inside the resulting rectangle.
// Pixels outside it are background (sequential/main.go:17, distributed/main.go:70) and are pre-cleared like every miss.
//
// Conservative by construction: the pixel of a point c is found by solving  c - cam = l * (fwd + left*si + up*sj)  for
// (l, si, sj) (no assumption that the basis is orthonormal), the box lies in front of the camera (l > 0 at all 8 corners, hence
// everywhere in the box), so its image is the convex hull of the 8 corner images; the rectangle is their bounding box grown by
// RT_CULL_PAD pixels (the fp32 primary direction of the fast kernel is within ~1e-6 of the exact one; a pixel is ~3e-4).  If any
// corner is not safely in front of the camera the object covers the whole screen as far as this test is concerned.
#pragma once
#include "rt_types.h"
#include <math.h>

#define RT_CULL_PAD 2

inline void rt_cull_full(FrameParams &fp) {
    fp.cull_on = 0; fp.cull_i0 = 0; fp.cull_j0 = 0; fp.cull_i1 = fp.W; fp.cull_j1 = fp.H;
    fp.rbx0 = fp.rby0 = fp.rnbx = fp.rnby = 0;
    fp.lt0 = fp.lt_n = 0;
}

// Fills the cull_* / r* fields of fp (whose camera, screen and region / tile fields are final).
inline void rt_cull_frame(FrameParams &fp, const MeshView *meshes, const ObjRec *objs, int n_obj) {
    rt_cull_full(fp);
    const double *f = fp.fwd, *l = fp.left, *u = fp.up;
    // rows of the inverse of [fwd left up] (times det)
    const double a[3] = { l[1] * u[2] - l[2] * u[1], l[2] * u[0] - l[0] * u[2], l[0] * u[1] - l[1] * u[0] };
    const double b[3] = { u[1] * f[2] - u[2] * f[1], u[2] * f[0] - u[0] * f[2], u[0] * f[1] - u[1] * f[0] };
    const double c[3] = { f[1] * l[2] - f[2] * l[1], f[2] * l[0] - f[0] * l[2], f[0] * l[1] - f[1] * l[0] };
    const double det = f[0] * a[0] + f[1] * a[1] + f[2] * a[2];
    const int hw = fp.W / 2, hh = fp.H / 2;
    const double phw = fp.tan_half, phh = phw * (double)fp.H / (double)fp.W;
    if (!(fabs(det) > 1e-12) || !(phw > 1e-12) || hw < 1 || hh < 1 || !isfinite(phw)) return;
    double imin = HUGE_VAL, imax = -HUGE_VAL, jmin = HUGE_VAL, jmax = -HUGE_VAL;
    bool any = false;
    for (int k = 0; k < n_obj; k++) {
        const MeshView &m = meshes[objs[k].mesh];
        if (m.n_faces == 0) continue;
        if (!(objs[k].pos[0] == objs[k].pos[0])) continue;        // NaN position: absent from this order, can never be hit
        for (int corner = 0; corner < 8; corner++) {
            double v[3];
            for (int d = 0; d < 3; d++)
                v[d] = (double)((corner >> d) & 1 ? m.hi[d] : m.lo[d]) + objs[k].pos[d] - fp.cam_pos[d];
            const double lam = (v[0] * a[0] + v[1] * a[1] + v[2] * a[2]) / det;
            const double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (!(lam > 1e-6 * (len + 1.0))) return;             // at or behind the camera plane (or NaN): no culling
            const double si = (v[0] * b[0] + v[1] * b[1] + v[2] * b[2]) / det / lam;
            const double sj = (v[0] * c[0] + v[1] * c[1] + v[2] * c[2]) / det / lam;
            // tracer.go:18-19 inverted:  si = phw*((hw - i) - 0.5)/hw
            const double pi = (double)hw - 0.5 - si * (double)hw / phw, pj = (double)hh - 0.5 - sj * (double)hh / phh;
            if (!isfinite(pi) || !isfinite(pj)) return;
            imin = fmin(imin, pi); imax = fmax(imax, pi); jmin = fmin(jmin, pj); jmax = fmax(jmax, pj);
        }
        any = true;
    }
    fp.cull_on = 1;
    if (!any) { fp.cull_i0 = fp.cull_i1 = fp.cull_j0 = fp.cull_j1 = 0; return; }          // nothing to hit at all
    const double big = 1e9;
    imin = fmax(imin, -big); jmin = fmax(jmin, -big); imax = fmin(imax, big); jmax = fmin(jmax, big);
    int i0 = (int)floor(imin) - RT_CULL_PAD, i1 = (int)ceil(imax) + RT_CULL_PAD + 1;
    int j0 = (int)floor(jmin) - RT_CULL_PAD, j1 = (int)ceil(jmax) + RT_CULL_PAD + 1;
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 > fp.W) i1 = fp.W;
    if (j1 > fp.H) j1 = fp.H;
    if (i1 <= i0 || j1 <= j0) { fp.cull_i0 = fp.cull_i1 = fp.cull_j0 = fp.cull_j1 = 0; return; }
    fp.cull_i0 = i0; fp.cull_i1 = i1; fp.cull_j0 = j0; fp.cull_j1 = j1;
    if (fp.tile_w > 0) {
        // tile mode: the kernels test every block against the rectangle (block_culled); here the launch is cut down to the tile
        // ROWS that meet it.  Tile ids of rows TY0..TY1 are the range [TY0*tiles_x, (TY1+1)*tiles_x); this rank's tiles
        // t = lt*n_ranks + rank inside it are the contiguous local range [lt_a, lt_b).
        const int y0 = j0 > fp.y ? j0 : fp.y, y1 = j1 < fp.y + fp.h ? j1 : fp.y + fp.h;
        if (y1 <= y0 || i1 <= fp.x || i0 >= fp.x + fp.w) { fp.cull_i0 = fp.cull_i1 = fp.cull_j0 = fp.cull_j1 = 0; return; }
        const long long ta = (long long)((y0 - fp.y) / fp.tile_h) * fp.tiles_x, tb = (long long)((y1 - 1 - fp.y) / fp.tile_h + 1) * fp.tiles_x;
        const long long n = fp.n_ranks, r = fp.rank;
        long long lt_a = ta > r ? (ta - r + n - 1) / n : 0, lt_b = tb > r ? (tb - r + n - 1) / n : 0;
        if (lt_b > fp.n_local_tiles) lt_b = fp.n_local_tiles;
        if (lt_b <= lt_a) { fp.cull_i0 = fp.cull_i1 = fp.cull_j0 = fp.cull_j1 = 0; return; }
        fp.lt0 = (int)lt_a; fp.lt_n = (int)(lt_b - lt_a);
        return;
    }
    // region mode: the block sub-rectangle of the region that intersects the pixel rectangle
    const int x0 = i0 > fp.x ? i0 : fp.x, x1 = i1 < fp.x + fp.w ? i1 : fp.x + fp.w;
    const int y0 = j0 > fp.y ? j0 : fp.y, y1 = j1 < fp.y + fp.h ? j1 : fp.y + fp.h;
    if (x1 <= x0 || y1 <= y0) return;                            // rnbx stays 0 with cull_on: ray_count() == 0
    fp.rbx0 = (x0 - fp.x) / 8; fp.rby0 = (y0 - fp.y) / 4;
    fp.rnbx = (x1 - 1 - fp.x) / 8 - fp.rbx0 + 1; fp.rnby = (y1 - 1 - fp.y) / 4 - fp.rby0 + 1;
}
