// rt_types.h -- records shared by the host code and the sm_100a kernels.
//
// HBM layout (per mesh, all built once at rt_scene_create; see DESIGN.md "Data layout"):
//   nodes     : wide-BVH node records (RT_W children: SoA boxes + refs; 128 B for RT_W = 4), BFS order
//   tri32     : FP32_FAST triangle records, 48 B (3 x float4), Morton/leaf order
//   tri64     : FP64_EXACT triangle records, 80 B (5 x double2), Morton/leaf order
//   fverts    : 9 doubles per face, ORIGINAL face order (winner re-intersection + flat normals)
//   fnormals  : 9 doubles per face, ORIGINAL face order (Phong normals), absent if the mesh has no vn
//   fmat      : material index per face; materials: 10 doubles each
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

#ifndef RT_W
#define RT_W 4                    // children per BVH node (compile-time; 4 measured best)
#endif
#define RT_STACK_SIZE 128         // entries of a ray's traversal stack (local memory; only the touched part costs anything)
#define RT_STACK_NODES (RT_STACK_SIZE - 3 * RT_W)   // the builder rejects trees with (RT_W-1)*depth + 2 > RT_STACK_NODES: the top
                                  // 3*RT_W entries are the least room the postponed leaves of an any-hit ray get (rt_traverse.cuh)
#define RT_LEAF_FIRST_BITS 27     // child ref < 0: ~ref = first | (count-1) << 27
#define RT_LEAF_FIRST_MASK ((1u << RT_LEAF_FIRST_BITS) - 1u)
#define RT_MAX_LEAF 16
#define RT_EMPTY 0x7fffffff       // unused child slot / "no node" sentinel (never a valid index)
#define RT_NODE_BYTES (28 * RT_W) // bytes fetched per node visit: 6 box arrays + refs, RT_W x 4 B each
// 256-bit loads (sm_100a: LDG.E.256).  The traversal is bound by the L1 data pipe (ncu: l1tex wavefronts 68 % of peak, the
// highest unit, while issue slots are half empty -- profiles/r02_ncu_c3_summary.txt), and a divergent warp pays one wavefront per
// lane and LOAD INSTRUCTION whatever its width: fetching a node with 3 x 32 B + 16 B instead of 7 x 16 B, and a triangle with
// 2 x 32 B (64-byte records) instead of 3 x 16 B, cuts the wavefronts of a ray by ~40 %.
#ifndef RT_LD256
#define RT_LD256 1
#endif
#ifndef RT_TRI_LD256
#define RT_TRI_LD256 RT_LD256                   // 64-byte triangle records fetched with 2 x LDG.256 (0: 48 bytes, 3 x LDG.128)
#endif
#define RT_TRI32_F4 (RT_TRI_LD256 ? 4 : 3)      // float4 slots per FP32_FAST triangle record
#define RT_TRI32_BYTES (16 * RT_TRI32_F4)
#define RT_TRI64_BYTES 80
#define RT_SHADOW_BIAS 0.0001     // tracer.go:64
// words of a lane's counter block (unsigned int[8]): [0] next ray (persistent K2), [1] hit count, [2] next hit (K4), [3] spare
#define RT_CNT_CANCEL 4           // cancel generation written by rt_cancel; a kernel gives up when it exceeds FrameParams.cancel_gen
#define RT_CNT_PUSH_DONE 5        // CTAs-done counter of k_push_tiles
// measured defaults (profiles/r02_kernel_ab.txt, C3 at N = 1): any-hit hierarchy with leaves of <= 2 triangles (K4 -2.8 %,
// +15 % node memory; single-triangle leaves are no faster and cost +93 %), fused frame kernel (-6.5 % pipelined, -12 % latency)
#ifndef RT_DEFAULT_MAX_LEAF_SHADE
#define RT_DEFAULT_MAX_LEAF_SHADE 2
#endif
// persistent SM-affine frame kernel (render.cu: k_frame_fused_sm): 8x4 blocks per chunk, per-SM words of the chunk state
#ifndef RT_CHUNK_BLOCKS
#define RT_CHUNK_BLOCKS 32
#endif
#define RT_MAX_SM 256
#ifndef RT_DEFAULT_FUSED
#define RT_DEFAULT_FUSED 1
#endif

struct alignas(16) rt_f4 { float x, y, z, w; };
struct alignas(32) rt_f8 { rt_f4 a, b; };
struct alignas(16) rt_i4 { int x, y, z, w; };

// One wide-BVH node: the boxes of its RT_W children in SoA form (each array is RT_W/4 16-byte vectors, fetched with
// LDG.128 through the read-only path) + the child refs.  ref >= 0: node index, ref < 0: leaf (see below),
// ref == RT_EMPTY: unused slot.  A ray visits a node once and tests all children with independent instructions, so the
// dependent chain per ray is the tree depth (log_W), ~2.2x shorter than with a binary tree for RT_W = 4.
struct alignas(32) RtNode {
    float lox[RT_W], hix[RT_W], loy[RT_W], hiy[RT_W], loz[RT_W], hiz[RT_W];
    int ref[RT_W];
    int pad[RT_W];
};

// EXPERIMENT, MEASURED SLOWER, off by default (-DRT_NODEQ=1; profiles/r02_kernel_ab.txt): the same RT_W = 4 children with their boxes
// on a 16-bit grid that spans the mesh (origin qorg, cell qstep per axis; MeshView), 64 bytes = 2 x LDG.E.256 instead of 112 bytes
// in 4 loads.  Motivation: ncu on the fused frame kernel shows the L1 data pipe as the busiest unit (69.6 % of peak) with 3.7
// data-pipe wavefronts per tag wavefront, while issue slots are 58 % used.  A plane is decoded with one PRMT (the 16-bit field
// becomes the mantissa of 2^23 + q) and folded into the slab test's FFMA: t = (2^23 + q) * (qstep / d) + ((qorg - o) / d -
// 2^23 * qstep / d).  Rounding: the folded constant costs at most half a cell, and every lower / upper plane is stored one whole
// cell further out (floor - 1 / ceil + 1 on a grid that starts two cells outside the mesh), so the quantised box always contains
// the float box it was made from: the hierarchy still only culls (identical frames, identical node counts per shadow ray).
// Result on C3: K4 0.555 -> 0.548 ms, K2 0.283 -> 0.323 ms, fused frame 0.603 -> 0.625 ms pipelined.  Halving the node bytes buys
// nothing and the 24 extra PRMT per visit cost the primary rays 14 %: the traversal is bound by neither L1 bytes nor DRAM but by
// the dependent latency of a visit at the occupancy 56 registers allow.
// word layout: x[0..3] = lo of children 0..3, x[4..7] = hi; likewise y, z; then the 4 child refs.
#ifndef RT_NODEQ
#define RT_NODEQ 0
#endif
struct alignas(32) RtNodeQ { uint16_t x[8], y[8], z[8]; int ref[4]; };
#define RT_QGRID_CELLS 65531.0f   // cells across the padded mesh extent; + 2 cells of margin on either side < 65536

RT_HD int rt_leaf_ref(uint32_t first, uint32_t count) { return ~(int)(first | ((count - 1u) << RT_LEAF_FIRST_BITS)); }
RT_HD uint32_t rt_leaf_first(int ref) { return ((uint32_t)~ref) & RT_LEAF_FIRST_MASK; }
RT_HD uint32_t rt_leaf_count(int ref) { return (((uint32_t)~ref) >> RT_LEAF_FIRST_BITS) + 1u; }

struct MeshView {
    const RtNode  *nodes;
    const rt_f4   *tri32;      // RT_TRI32_F4 per triangle: (P1,faceidx) (P2,-) (P3,-) [(pad)]
    const double  *tri64;      // 10 per triangle: P1 P2 P3, [9] = face index (as double, exact)
    const double  *fverts;     // 9 per face
    const double  *fnormals;   // 9 per face or nullptr
    const uint32_t *fmat;
    const double  *materials;  // 10 per material
    uint32_t n_faces, n_nodes;
    int      root_ref;         // ref encoding; meaningful iff n_faces > 0
    float    lo[3], hi[3];     // padded mesh bounds (object space)
    // second hierarchy over the SAME triangle array for any-hit (shadow) rays: its leaves may be smaller (a warp pays for its
    // largest leaf; RT_B200_MAX_LEAF_SHADE).  Equal to nodes / root_ref when both leaf sizes agree.
    const RtNode  *nodes_any;
    int      root_ref_any;
    uint32_t n_nodes_any;
    uint32_t any_defer;                   // any-hit rays test their postponed leaves once this many are waiting (rt_traverse.cuh); >= 1
    const RtNodeQ *qnodes, *qnodes_any;   // FP32_FAST: the 64-byte quantised form of nodes / nodes_any (null in FP64_EXACT scenes)
    float    qorg[3], qstep[3];           // its grid: plane = qorg + q * qstep
    double   vlo[3], vhi[3];   // float64 bounds of ALL vertices of the mesh: Object.Bounds (object.go:38-49) for tracer.go:32's test
};

struct ObjRec { double pos[3]; uint32_t mesh; uint32_t pad; };
struct LightRec { double pos[3]; double col[3]; };

struct FrameParams {
    double cam_pos[3], fwd[3], left[3], up[3];
    double tan_half;
    int W, H;            // full screen (tracer.go:81)
    int x, y, w, h;      // region     (distributed/main.go:50-51)
    int n_lights, n_obj;
    // tile mode (multi-GPU): if tile_w > 0 the region-linear pixel index maps through interleaved tiles
    int tile_w, tile_h, tiles_x, rank, n_ranks, n_local_tiles;
    int tile_final;      // tile mode: output index = position in the FULL framebuffer (peer-memory writer), not in the packed tiles
    // Primary-ray culling (rt_cull.h): no primary ray outside the absolute pixel rectangle [cull_i0, cull_i1) x [cull_j0, cull_j1)
    // can meet an object's root box.  Region mode traces only the 8x4 blocks rbx0 .. rbx0+rnbx-1, rby0 .. rby0+rnby-1 (relative to
    // the region origin); rnbx == 0 means "every block of the region" (no culling information).  Tile mode keeps its ray
    // numbering and skips the blocks that lie outside the rectangle (cull_on).
    int cull_on, cull_i0, cull_j0, cull_i1, cull_j1;
    int rbx0, rby0, rnbx, rnby;
    int lt0, lt_n;
    // ^ tile mode: only the local tiles lt0 .. lt0+lt_n-1 are traced (the tile ROWS that meet the rectangle: the local tiles of a
    // range of rows are a contiguous range of local indices); lt_n == 0 means all of them
    unsigned int cancel_gen;   // generation of the call (rt_cancel): the kernels stop when counters[RT_CNT_CANCEL] has moved past it
};

// Per-frame mutable state small enough to travel BY VALUE as a __grid_constant__ kernel parameter: no device
// upload, no host synchronisation per frame (the gob `diff` of a WorkOrder is < 1 KB: distributed/main.go:56-64).
#define RT_INLINE_LIGHTS 8
#define RT_INLINE_OBJS 4
struct InlineFrame { LightRec lights[RT_INLINE_LIGHTS]; ObjRec objs[RT_INLINE_OBJS]; int use_lights, use_objs; };

// primary-hit record appended by the primary kernel, consumed by the shade kernel (16 B)
struct HitRec { uint32_t pix; int32_t obj; int32_t face; float t; };

struct Counters { unsigned long long nodes, tris; };

// frame hand-off record in peer memory (one per framebuffer slot, RT_SYNC_SLOT_BYTES = 256 in include/rt_b200.h), uint32 words:
#define RT_SYNC_ARRIVE 0        // [0..31]  arrive[rank]
#define RT_SYNC_CONSUMED 32
#define RT_SYNC_ERROR 33
struct PushSync { uint32_t *slot; unsigned int *blocks_done; uint32_t seq, timeout_ms, debug; };

// ------------------------------------------------------------------------------------------------
// ray index <-> pixel mapping.  Rays are numbered so that 32 consecutive indices form an 8x4 pixel
// block (coherent warps).  Region mode: blocks row-major over the region.  Tile mode (multi-GPU):
// tiles t = lt*n_ranks + rank at screen position tile_pos(t), each tile_w x tile_h (multiples of 8 x 4), blocks row-major in a tile.
// ------------------------------------------------------------------------------------------------
struct PixMap { int i, j; uint32_t out; bool valid; };

// Tile id t (owner rank = t % n_ranks, slot in the owner's packed buffer = t / n_ranks)  <->  tile position (X, Y) on the
// screen.  Tile row Y is rotated by 3*Y positions: when n_ranks divides tiles_x (3840/32 = 120 = 8*15) the plain row-major
// numbering would give every rank whole tile COLUMNS (vertical stripes; 9.6 % more hit pixels on the busiest of 8 ranks for
// the C3 view); rotated, the owners form diagonals and the imbalance is < 1 % for 2..8 ranks.
RT_HD void tile_pos(uint32_t t, uint32_t tiles_x, uint32_t &X, uint32_t &Y) {
    Y = t / tiles_x;
    X = (t % tiles_x + 3u * Y) % tiles_x;
}
RT_HD uint32_t tile_id(uint32_t X, uint32_t Y, uint32_t tiles_x) {
    return Y * tiles_x + (X + tiles_x - (3u * Y) % tiles_x) % tiles_x;
}   // i,j = absolute screen pixel; out = index in the output buffer

RT_HD uint32_t ray_count(const FrameParams &fp) {
    if (fp.tile_w > 0) {
        if (fp.cull_on && fp.cull_i1 <= fp.cull_i0) return 0u;     // nothing can be hit
        return (uint32_t)(fp.lt_n > 0 ? fp.lt_n : fp.n_local_tiles) * (uint32_t)(fp.tile_w * fp.tile_h);
    }
    if (fp.rnbx > 0) return (uint32_t)fp.rnbx * (uint32_t)fp.rnby * 32u;
    if (fp.cull_on) return 0u;                                   // culling found no block that can be hit
    uint32_t nbx = (uint32_t)(fp.w + 7) / 8, nby = (uint32_t)(fp.h + 3) / 4;
    return nbx * nby * 32u;
}

// tile mode: can the 8x4 block whose first pixel is (i0, j0) be skipped?  (i0, j0) = pixel of the ray minus its lane offsets
// (lane & 7, lane >> 3): blocks are aligned to the REGION origin, not to the screen
RT_HD bool block_culled(const FrameParams &fp, int i0, int j0) {
    return fp.cull_on && (i0 + 8 <= fp.cull_i0 || i0 >= fp.cull_i1 || j0 + 4 <= fp.cull_j0 || j0 >= fp.cull_j1);
}

RT_HD PixMap map_ray(uint32_t r, const FrameParams &fp, bool rowmajor) {
    PixMap p;
    uint32_t lane = r & 31u, blk = r >> 5;
    int lx = (int)(lane & 7u), ly = (int)(lane >> 3);
    if (fp.tile_w > 0) {
        // tiles are cut from the REGION (x, y, w, h); a full-screen frame is the region (0, 0, W, H)
        uint32_t bpt = (uint32_t)(fp.tile_w * fp.tile_h) / 32u;       // blocks per tile
        uint32_t lt = (uint32_t)fp.lt0 + blk / bpt, b = blk % bpt;
        uint32_t bx = b % (uint32_t)(fp.tile_w / 8), by = b / (uint32_t)(fp.tile_w / 8);
        int tx = (int)bx * 8 + lx, ty = (int)by * 4 + ly;
        uint32_t t = lt * (uint32_t)fp.n_ranks + (uint32_t)fp.rank;
        uint32_t TX, TY;
        tile_pos(t, (uint32_t)fp.tiles_x, TX, TY);
        const int ri = (int)TX * fp.tile_w + tx, rj = (int)TY * fp.tile_h + ty;
        p.i = fp.x + ri;
        p.j = fp.y + rj;
        p.out = lt * (uint32_t)(fp.tile_w * fp.tile_h) + (uint32_t)(ty * fp.tile_w + tx);
        p.valid = ri < fp.w && rj < fp.h;
        if (fp.tile_final && p.valid)
            p.out = rowmajor ? (uint32_t)rj * (uint32_t)fp.w + (uint32_t)ri : (uint32_t)ri * (uint32_t)fp.h + (uint32_t)rj;
        return p;
    }
    uint32_t nbx = fp.rnbx > 0 ? (uint32_t)fp.rnbx : (uint32_t)(fp.w + 7) / 8;
    int ri = ((int)(blk % nbx) + fp.rbx0) * 8 + lx, rj = ((int)(blk / nbx) + fp.rby0) * 4 + ly;
    p.valid = ri < fp.w && rj < fp.h;
    p.i = fp.x + ri; p.j = fp.y + rj;
    p.out = rowmajor ? (uint32_t)rj * (uint32_t)fp.w + (uint32_t)ri        // image order
                     : (uint32_t)ri * (uint32_t)fp.h + (uint32_t)rj;       // distributed/main.go:82
    return p;
}
