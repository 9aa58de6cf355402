/*
 * rt_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C float64 restatement of the render hot path of
 * MWindels/distributed-raytracer (Go).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or images,
 * cannot be built here (no Go toolchain; rtreego / gwob / go-sdl2 absent),
 * so this restatement is pinned only by (1) the reference source it follows
 * line by line (citations below) and (2) the known-answer values in
 * SURVEY.md Appendix C, produced by an independent Python restatement, and
 * (3) tests/pyref.py, a second restatement written from the Go sources alone,
 * which agrees with this one bit for bit (tests/test_oracle.py).  The one
 * pinned function is orc_encode_trace_results (checked against the
 * google.protobuf runtime, tests/test_wire.py).
 *
 * Modes: the R-tree mode is the reference-shaped one (candidates = faces whose rectangle the
 * ray pierces, object.go:76).  Brute force over all faces gives the same result EXCEPT for
 * zero-area faces, whose `incidence` (triangle.go:42) is a rounding residue and which can then
 * "hit" at t ~ 0 for rays nowhere near them (profiles/r01_fuzz_cpu_replay.txt).
 *
 * All citations are into /root/reference/.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Go math library restatements (Go stdlib, math/tan.go, math/pow.go) -- */
double orc_go_tan(double x);
double orc_go_pow(double x, double y);
double orc_go_max(double x, double y);   /* math.Max, Go stdlib math/dim.go */
double orc_go_min(double x, double y);   /* math.Min */

/* ---- shared/geom ---------------------------------------------------------- */
/* triangle.go:37-77.  tri = P1,P2,P3 (9 doubles).  Returns 1 on hit and
 * fills ip[3] (o + d*t), bary[3] (r1,r2,r3) and *t_out (dirScale). */
int orc_triangle_intersection(const double *tri, const double *o, const double *d,
                              double *ip, double *bary, double *t_out);
/* box.go:29-69 */
int orc_box_intersect(const double *bmin, const double *bmax, const double *o, const double *d);

/* ---- shared/state/camera.go:35-44 ---------------------------------------- */
/* out basis = forward[3], left[3], up[3].  Returns -1 if dir x GlobalUp == 0. */
int orc_new_camera(const double *dir, double *basis9);

/* tracer.go:15-22,86 : primary ray direction for pixel (i,j) of a W x H screen.
 * cam13 = pos[3], forward[3], left[3], up[3], fov. */
void orc_primary_ray(int i, int j, int W, int H, const double *cam13, double *dir3);

/* ---- scene ---------------------------------------------------------------- */
typedef struct orc_scene orc_scene;

typedef struct {
    const double   *verts;      /* nv*3, object space (mesh.go:102)                       */
    uint32_t        n_verts;
    const double   *normals;    /* nn*3, normalised at load (mesh.go:202); nn may be 0    */
    uint32_t        n_normals;
    const uint32_t *faces;      /* nf*7 : v0 v1 v2 n0 n1 n2 mat   (mesh.go:21-27)         */
    uint32_t        n_faces;
    const double   *materials;  /* nm*10: Ka[3] Kd[3] Ks[3] Ns    (mesh.go:94-97)         */
    uint32_t        n_materials;
} orc_mesh_desc;

/* mode for candidate enumeration (results are identical; speed differs) */
enum {
    ORC_BRUTE   = 0, /* every face in index order, strict '<'  (object.go:76-105 w/o culling) */
    ORC_CULLED  = 1, /* conservative median-split BVH cull, winner = lexicographic (dist,index) */
    ORC_RTREE   = 2  /* reference-shaped: Guttman R-tree (3,2,5) + box.go predicate, no tmax   */
};

orc_scene *orc_scene_create(const orc_mesh_desc *meshes, uint32_t n_meshes,
                            const double *obj_pos /* n_obj*3 */, const uint32_t *obj_mesh,
                            uint32_t n_obj);
void orc_scene_destroy(orc_scene *);
/* build the reference-shaped R-trees (mesh.go:139,208) now instead of lazily */
void orc_build_rtrees(orc_scene *);
int orc_rtree_path_check(const orc_scene *s, uint32_t mesh, int32_t face, const double *o, const double *d, int *first_rejected,
                         double *bmin, double *bmax);

/* per-frame mutable state (environment.go:65-69) */
void orc_frame_set(orc_scene *, const double *cam13, double tan_half_fov,
                   const double *lights /* n*6: pos[3] col[3] */, uint32_t n_lights,
                   const double *obj_pos /* may be NULL = keep */);

/* tracer.go:27-50 : nearest hit.  Returns 1 on hit.
 * out: ip[3], normal[3], *face (index within its mesh), *obj, *mat (index in mesh). */
int orc_trace(const orc_scene *, int mode, const double *o, const double *d,
              double *ip, double *normal, int32_t *face, int32_t *obj, int32_t *mat);

/* tracer.go:81-91 : one pixel.  Returns hit flag; colour[3] fp64; face/obj of the
 * primary hit; lit_mask bit k = light k contributed (tracer.go:64). */
int orc_trace_pixel(const orc_scene *, int mode, int i, int j, int W, int H,
                    double *colour, int32_t *face, int32_t *obj, uint32_t *lit_mask);

/* quantisers: 0 = colour.go:59-61 RGB() uint8(255*c); 1 = go-sdl2 path uint32(0xFFFF*c)>>8 */
enum { ORC_Q_RGB255 = 0, ORC_Q_RGBA16_SHR8 = 1 };
/* layouts: 0 = BulkTrace column-major region index i*h+j (distributed/main.go:82);
 *          1 = row-major j*w+i */
enum { ORC_L_COLMAJOR = 0, ORC_L_ROWMAJOR = 1 };

typedef struct {
    uint64_t primary_rays, shadow_rays, hit_pixels;
    uint64_t tri_tests, box_tests;
    uint64_t nearest_ties;      /* exact-distance ties seen while selecting (Appendix B-1) */
} orc_stats;

/* distributed/main.go:67-88 / sequential/main.go:21-28 : region render.
 * out_rgb   : w*h*3 bytes (may be NULL)
 * out_colour: w*h*3 doubles (may be NULL), out_face: w*h int32 (may be NULL; -1 = miss),
 * all in the chosen layout.  step>1 renders only pixels whose region-linear index
 * (i*h+j) % step == phase (bounded sample for CPU-baseline timing); others untouched.
 * n_threads<=0: all cores.  Returns 0. */
int orc_render_region(const orc_scene *, int mode, int W, int H, int x, int y, int w, int h,
                      int layout, int quant, int n_threads, int step, int phase,
                      uint8_t *out_rgb, double *out_colour, int32_t *out_face,
                      orc_stats *stats);

/* Result wire format: the bytes proto.Marshal emits for comms.TraceResults
 * (shared/comms/comms.proto:34-41), the message BulkTrace fills one Colour per pixel
 * (worker/distributed/main.go:52-54,79-86).  protobuf (third-party, unpinned in the reference; the
 * encoding itself is the published proto3 wire format): every element of `repeated Colour results = 1`
 * is tag 0x0A + varint length + payload; payload = fields r,g,b (tags 0x08,0x10,0x18) in number
 * order, a zero value is omitted, the value is a base-128 varint.  PINNED: tests/test_wire.py
 * compares this function with the google.protobuf runtime on the same message (and with the
 * committed vectors tests/golden/trace_results_*.bin that runtime produced).
 * rgb: n*3 bytes; out: at least 11*n bytes (may be NULL to only measure).  Returns the length. */
uint64_t orc_encode_trace_results(const uint8_t *rgb, uint64_t n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
