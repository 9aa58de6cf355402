// bvh_build.cu -- K1: on-GPU LBVH build (sm_100a).
//
// Replaces rtreego.NewTree(3,2,5) + one serial Insert per face
// (/root/reference/shared/state/mesh.go:139,208).  Pipeline, all on the device:
//   k_mesh_bounds   face vertices -> mesh bounds (warp-reduced ordered-int atomics)
//   k_leaf_morton   per face: padded fp32 leaf box + 63-bit Morton key of the box centroid
//   cub radix sort  (key, face) pairs                                  [library plumbing]
//   k_karras        binary radix tree over the sorted keys (Karras 2012), parent links
//   k_refit_sah     bottom-up boxes + SAH cost, collapse subtrees of <= max_leaf faces into leaves
//   k_emit_count_d / scan / k_emit_write_d / k_emit_advance   BFS re-emission into RT_W-wide traversal nodes (greedy
//                   surface-area collapse), frontier size kept on the device: one host round trip per 8 levels
//   k_gather_tris   triangle records in leaf order: 48 B fp32 and/or 80 B fp64
#include "rt_internal.h"
#include <cub/cub.cuh>

namespace {

__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned int u) {
    unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// bounds[0..2] = min (ordered uint), bounds[3..5] = max
__global__ void k_mesh_bounds(const double *__restrict__ fverts, uint32_t n, unsigned int *bounds) {
    float lo[3] = { HUGE_VALF, HUGE_VALF, HUGE_VALF }, hi[3] = { -HUGE_VALF, -HUGE_VALF, -HUGE_VALF };
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        const double *p = fverts + 9 * (size_t)f;
        for (int v = 0; v < 3; v++)
            for (int k = 0; k < 3; k++) {
                double c = p[3 * v + k];
                lo[k] = fminf(lo[k], f32_down(c));
                hi[k] = fmaxf(hi[k], f32_up(c));
            }
    }
    for (int k = 0; k < 3; k++) {
        unsigned int l = f2ord(lo[k]), h = f2ord(hi[k]);
        l = __reduce_min_sync(0xffffffffu, l);
        h = __reduce_max_sync(0xffffffffu, h);
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], l); atomicMax(&bounds[3 + k], h); }
    }
}

__global__ void k_leaf_morton(const double *__restrict__ fverts, uint32_t n, BuildParams bp, BoxF *__restrict__ leafbox,
                              unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) leaf_morton_thread(f, fverts, bp, leafbox, keys, vals);
}

__global__ void k_karras(const unsigned long long *__restrict__ keys, int n, KarrasNode *__restrict__ kn,
                         int *__restrict__ parent_int, int *__restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) karras_thread(i, keys, n, kn, parent_int, parent_leaf);
}

__global__ void k_refit_sah(int n, const uint32_t *__restrict__ vals, const BoxF *__restrict__ leafbox,
                            const KarrasNode *__restrict__ kn, const int *__restrict__ parent_int,
                            const int *__restrict__ parent_leaf, BoxF *ibox, float *icost, unsigned char *collapsed,
                            unsigned int *flags, int max_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) refit_sah_thread(i, vals, leafbox, kn, parent_int, parent_leaf, ibox, icost, collapsed, flags, max_leaf);
}

// ---- BFS emission with the frontier size on the DEVICE ------------------------------------------------------------
// state[0] = frontier size of the current level, [1] = index of its first node, [2] = levels emitted, [3] = total leaf refs.
// Every kernel of a level is launched for the widest possible frontier and reads the real size from `state`, so a batch of
// levels is enqueued without a host round trip (rt_build_bvh reads the state back once per RT_EMIT_BATCH levels).
#define RT_EMIT_BATCH 8
__global__ void k_emit_count_d(const int *__restrict__ frontier, const int *__restrict__ state, int n_max, const KarrasNode *__restrict__ kn,
                               const unsigned char *__restrict__ collapsed, const BoxF *__restrict__ ibox, int *__restrict__ cnt) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_max) return;
    if (p < state[0]) emit_count_thread(p, frontier, kn, collapsed, ibox, cnt);
    else cnt[p] = 0;                                              // the scan runs over n_max entries
}
__global__ void k_emit_write_d(const int *__restrict__ frontier, const int *__restrict__ state, const KarrasNode *__restrict__ kn,
                               const unsigned char *__restrict__ collapsed, const int *__restrict__ off, const uint32_t *__restrict__ vals,
                               const BoxF *__restrict__ leafbox, const BoxF *__restrict__ ibox, RtNode *__restrict__ nodes,
                               int *__restrict__ next_frontier, unsigned int *__restrict__ leaf_counter) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int fsize = state[0];
    if (p >= fsize) return;
    unsigned int nleaf = emit_write_thread(p, frontier, fsize, state[1], kn, collapsed, off, vals, leafbox, ibox, nodes, next_frontier);
    if (nleaf) atomicAdd(leaf_counter, nleaf);
}
// next level: base += fsize; fsize = number of internal children just emitted
__global__ void k_emit_advance(int *state, const int *__restrict__ cnt, const int *__restrict__ off) {
    const int fsize = state[0];
    if (fsize <= 0) return;
    state[1] += fsize;
    state[0] = off[fsize - 1] + cnt[fsize - 1];
    state[2] += 1;
}

__global__ void k_gather_tris(const double *__restrict__ fverts, const uint32_t *__restrict__ vals, uint32_t n,
                              rt_f4 *__restrict__ tri32, double *__restrict__ tri64) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gather_tri_thread(i, fverts, vals, tri32, tri64);
}

// min / max of all vertex coordinates of a mesh in float64 (Object.Bounds, object.go:38-49), through an order-preserving map to
// unsigned integers (-0 sorts below +0, as Go's math.Min / math.Max have it)
__device__ __forceinline__ unsigned long long d2ord(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static double ord2d(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double v; memcpy(&v, &u, 8); return v;
}
__global__ void k_vertex_bounds(const double *__restrict__ verts, uint32_t n_verts, unsigned long long *bounds) {
    unsigned long long lo[3] = { ~0ull, ~0ull, ~0ull }, hi[3] = { 0ull, 0ull, 0ull };
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_verts; v += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; k++) {
            const unsigned long long o = d2ord(verts[3 * (size_t)v + k]);
            lo[k] = o < lo[k] ? o : lo[k];
            hi[k] = o > hi[k] ? o : hi[k];
        }
    for (int k = 0; k < 3; k++) {
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long a = __shfl_down_sync(0xffffffffu, lo[k], s), b = __shfl_down_sync(0xffffffffu, hi[k], s);
            lo[k] = a < lo[k] ? a : lo[k];
            hi[k] = b > hi[k] ? b : hi[k];
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], lo[k]); atomicMax(&bounds[3 + k], hi[k]); }
    }
}

__global__ void k_quantise_nodes(const RtNode *__restrict__ nodes, uint32_t n, RtNodeQ *__restrict__ out, float ox, float oy, float oz,
                                 float sx, float sy, float sz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const float qorg[3] = { ox, oy, oz }, qstep[3] = { sx, sy, sz };
    if (i < n) quantise_node_thread(i, nodes, out, qorg, qstep);
}

// scratch allocations of one build, released on scope exit
struct Scratch {
    std::vector<void *> ptrs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    MeshDevice *m = nullptr;
    bool ok = false;
    // one arena for the ~25 temporary arrays of a build: cudaMalloc / cudaFree each synchronise the device and cost more than the
    // kernels they serve (the build kernels of a 1 M-face mesh run ~3 ms; 25 allocations + frees made it ~40 ms)
    // (Taking the arena from the stream-ordered pool instead was measured and makes no difference: on a quiet box rt_scene_create of
    // the 991 232-face mesh takes 18 ms either way, both BVH builds 4.4 ms; the 20 ... 900 ms that fresh boxes sometimes show are
    // the driver mapping memory for the first allocations on a busy host.)
    char *arena = nullptr; size_t arena_size = 0, arena_used = 0;
    void reserve(size_t bytes) {
        if (cudaMalloc((void **)&arena, bytes) == cudaSuccess) { arena_size = bytes; ptrs.push_back(arena); }
        else { arena = nullptr; cudaGetLastError(); }                   // no room for the arena: fall back to one allocation per array
    }
    template <class T> cudaError_t alloc(T **p, size_t bytes) {
        bytes = (bytes ? bytes : 16);
        const size_t at = (arena_used + 255) & ~(size_t)255;
        if (arena && at + bytes <= arena_size) { *p = (T *)(arena + at); arena_used = at + bytes; return cudaSuccess; }
        cudaError_t e = cudaMalloc((void **)p, bytes);
        if (e == cudaSuccess) ptrs.push_back((void *)*p);
        return e;
    }
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (!ok && m) {
            if (m->nodes_any != m->nodes) cudaFree(m->nodes_any);
            if (m->qnodes_any != m->qnodes) cudaFree(m->qnodes_any);
            cudaFree(m->qnodes);
            cudaFree(m->nodes); cudaFree(m->tri32); cudaFree(m->tri64);
            m->nodes = nullptr; m->nodes_any = nullptr; m->tri32 = nullptr; m->tri64 = nullptr; m->qnodes = nullptr; m->qnodes_any = nullptr;
        }
    }
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

}  // namespace

// Builds nodes / tri32 / tri64 of `m` from m.fverts (device, 9 doubles per face).  Returns false + err.
// K0: the SoA hand-off.  The caller's INDEXED mesh (verts, normals, 7 indices per face: what
// state.Mesh holds, shared/state/mesh.go:21-27,100-106) is uploaded as it is and gathered into the per-face records on the
// device -- one thread per face; indices are validated here too (err: 1 vertex, 2 normal, 3 material index out of range).
__global__ void k_flatten_faces(const double *__restrict__ verts, const double *__restrict__ normals, const uint32_t *__restrict__ faces,
                                uint32_t n_faces, uint32_t n_verts, uint32_t n_normals, uint32_t n_materials, double *__restrict__ fverts,
                                double *__restrict__ fnormals, uint32_t *__restrict__ fmat, unsigned int *err) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const uint32_t *fc = faces + 7 * (size_t)f;
    unsigned int bad = 0;
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const uint32_t iv = fc[v];
        if (iv >= n_verts) { bad = 1; continue; }
#pragma unroll
        for (int k = 0; k < 3; k++) fverts[9 * (size_t)f + 3 * v + k] = verts[3 * (size_t)iv + k];
        if (n_normals) {
            const uint32_t in = fc[3 + v];
            if (in >= n_normals) { bad = bad ? bad : 2; continue; }
#pragma unroll
            for (int k = 0; k < 3; k++) fnormals[9 * (size_t)f + 3 * v + k] = normals[3 * (size_t)in + k];
        }
    }
    if (fc[6] >= n_materials) bad = bad ? bad : 3;
    else fmat[f] = fc[6];
    if (bad) atomicMax(err, 4u - bad);            // report the first kind in the order vertex, normal, material
}

// d_verts / d_normals / d_faces: the uploaded indexed mesh (device).  Allocates and fills m.fverts / m.fnormals / m.fmat.
// *bad: 0 ok, 1 vertex, 2 normal, 3 material index out of range.
cudaError_t rt_flatten_faces(MeshDevice &m, const double *d_verts, uint32_t n_verts, const double *d_normals, uint32_t n_normals,
                             const uint32_t *d_faces, cudaStream_t st, int *bad) {
    *bad = 0;
    const size_t nf = m.n_faces;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&m.fverts, sizeof(double) * 9 * (nf ? nf : 1))) != cudaSuccess) return e;
    if (n_normals && (e = cudaMalloc((void **)&m.fnormals, sizeof(double) * 9 * (nf ? nf : 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&m.fmat, sizeof(uint32_t) * (nf ? nf : 1))) != cudaSuccess) return e;
    if (!nf) return cudaSuccess;
    unsigned int *d_err = nullptr;
    if ((e = cudaMalloc((void **)&d_err, 8 + 6 * sizeof(unsigned long long))) != cudaSuccess) return e;
    cudaMemsetAsync(d_err, 0, sizeof(unsigned int), st);
    unsigned long long *d_vb = (unsigned long long *)(d_err + 2), h_vb[6] = { ~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull };
    cudaMemcpyAsync(d_vb, h_vb, sizeof h_vb, cudaMemcpyHostToDevice, st);
    if (n_verts) k_vertex_bounds<<<(unsigned)std::min<size_t>(((size_t)n_verts + 255) / 256, 1024), 256, 0, st>>>(d_verts, n_verts, d_vb);
    k_flatten_faces<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(d_verts, d_normals, d_faces, (uint32_t)nf, n_verts, n_normals, m.n_materials,
                                                                  m.fverts, m.fnormals, m.fmat, d_err);
    unsigned int h_err = 0;
    e = cudaMemcpyAsync(&h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_vb, d_vb, sizeof h_vb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_err);
    for (int k = 0; k < 3; k++) { m.vlo[k] = n_verts ? ord2d(h_vb[k]) : HUGE_VAL; m.vhi[k] = n_verts ? ord2d(h_vb[3 + k]) : -HUGE_VAL; }
    if (e == cudaSuccess && h_err) *bad = 4 - (int)h_err;
    return e;
}

bool rt_build_bvh(MeshDevice &m, bool want32, bool want64, int max_leaf, int max_leaf_any, int sm_count, cudaStream_t st, std::string &err) {
    const uint32_t n = m.n_faces;
    m.nodes = nullptr; m.nodes_any = nullptr; m.tri32 = nullptr; m.tri64 = nullptr; m.n_nodes = 0; m.n_nodes_any = 0; m.root_ref = 0; m.root_ref_any = 0;
    m.qnodes = nullptr; m.qnodes_any = nullptr;
    memset(&m.info, 0, sizeof m.info);
    m.info.n_faces = n; m.info.max_leaf = (uint32_t)max_leaf;
    for (int k = 0; k < 3; k++) { m.lo[k] = 0.f; m.hi[k] = 0.f; }
    if (n == 0) return true;
    if (n > RT_LEAF_FIRST_MASK) { err = "mesh has too many faces for the 27-bit leaf encoding"; return false; }
    if (max_leaf < 1) max_leaf = 1;
    if (max_leaf > RT_MAX_LEAF) max_leaf = RT_MAX_LEAF;
    if (max_leaf_any < 1) max_leaf_any = max_leaf;
    if (max_leaf_any > RT_MAX_LEAF) max_leaf_any = RT_MAX_LEAF;
    m.info.max_leaf_any = (uint32_t)max_leaf_any;

    unsigned int *d_bounds = nullptr; BoxF *d_leafbox = nullptr, *d_ibox = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr; uint32_t *d_vals = nullptr, *d_vals2 = nullptr;
    KarrasNode *d_kn = nullptr; int *d_pint = nullptr, *d_pleaf = nullptr; float *d_icost = nullptr;
    unsigned char *d_coll = nullptr; unsigned int *d_flags = nullptr, *d_leafcnt = nullptr;
    int *d_front[2] = { nullptr, nullptr }, *d_cnt = nullptr, *d_off = nullptr, *d_state = nullptr;
    void *d_tmp = nullptr; size_t tmp_bytes = 0, tmp_scan = 0;
    Scratch sc; sc.m = &m;
    const int TB = 256;
    const unsigned int nb = (n + TB - 1) / TB;
    unsigned int hb[6];
    BuildParams bp;
    uint32_t *sorted_vals = nullptr;
    unsigned long long *sorted_keys = nullptr;

    CK(cudaEventCreate(&sc.ev0)); CK(cudaEventCreate(&sc.ev1));
    CK(cudaEventRecord(sc.ev0, st));
    sc.reserve((size_t)n * 176 + ((size_t)48 << 20));       // ~150 bytes of temporaries per face + the CUB work space; overflow falls back
    CK(sc.alloc(&d_bounds, 6 * sizeof(unsigned int)));
    {
        unsigned int init[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
        CK(cudaMemcpyAsync(d_bounds, init, sizeof init, cudaMemcpyHostToDevice, st));
    }
    k_mesh_bounds<<<min(nb, (unsigned)(sm_count > 0 ? sm_count : 1) * 8u), TB, 0, st>>>(m.fverts, n, d_bounds);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hb, d_bounds, sizeof hb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = ord2f(hb[k]); hi[k] = ord2f(hb[3 + k]); }
        bp = make_build_params(lo, hi, m.lo, m.hi);
    }
    CK(sc.alloc(&d_leafbox, sizeof(BoxF) * (size_t)n));
    CK(sc.alloc(&d_keys, sizeof(unsigned long long) * (size_t)n));
    CK(sc.alloc(&d_keys2, sizeof(unsigned long long) * (size_t)n));
    CK(sc.alloc(&d_vals, sizeof(uint32_t) * (size_t)n));
    CK(sc.alloc(&d_vals2, sizeof(uint32_t) * (size_t)n));
    k_leaf_morton<<<nb, TB, 0, st>>>(m.fverts, n, bp, d_leafbox, d_keys, d_vals);
    CK(cudaGetLastError());
    {
        cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys2);
        cub::DoubleBuffer<uint32_t> vb(d_vals, d_vals2);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 63, st));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (int *)nullptr, (int *)nullptr, (int)n, st));
        if (tmp_scan > tmp_bytes) tmp_bytes = tmp_scan;
        CK(sc.alloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
        size_t tb = tmp_bytes;
        CK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, (int)n, 0, 63, st));
        sorted_keys = kb.Current(); sorted_vals = vb.Current();
    }
    if (want32) CK(cudaMalloc(&m.tri32, sizeof(rt_f4) * RT_TRI32_F4 * (size_t)n));
    if (want64) CK(cudaMalloc(&m.tri64, sizeof(double) * 10 * (size_t)n));
    k_gather_tris<<<nb, TB, 0, st>>>(m.fverts, sorted_vals, n, m.tri32, m.tri64);
    CK(cudaGetLastError());

    if (n == 1) {
        m.root_ref = m.root_ref_any = rt_leaf_ref(0, 1);
        m.info.n_leaves = 1; m.info.max_depth = 0;
    } else {
        const int ni = (int)n - 1;
        const unsigned int nbi = (ni + TB - 1) / TB;
        CK(sc.alloc(&d_kn, sizeof(KarrasNode) * (size_t)ni));
        CK(sc.alloc(&d_pint, sizeof(int) * (size_t)ni));
        CK(sc.alloc(&d_pleaf, sizeof(int) * (size_t)n));
        CK(sc.alloc(&d_ibox, sizeof(BoxF) * (size_t)ni));
        CK(sc.alloc(&d_icost, sizeof(float) * (size_t)ni));
        CK(sc.alloc(&d_coll, (size_t)ni));
        CK(sc.alloc(&d_flags, sizeof(unsigned int) * (size_t)ni));
        CK(sc.alloc(&d_leafcnt, sizeof(unsigned int)));
        CK(sc.alloc(&d_front[0], sizeof(int) * (size_t)ni));
        CK(sc.alloc(&d_front[1], sizeof(int) * (size_t)ni));
        CK(sc.alloc(&d_cnt, sizeof(int) * (size_t)(ni + 1)));
        CK(sc.alloc(&d_off, sizeof(int) * (size_t)(ni + 1)));
        CK(sc.alloc(&d_state, 4 * sizeof(int)));
        k_karras<<<nbi, TB, 0, st>>>(sorted_keys, (int)n, d_kn, d_pint, d_pleaf);
        CK(cudaGetLastError());
        // One wide hierarchy per leaf size over the one binary radix tree: refit + SAH collapse decision, then BFS emission.
        auto emit = [&](int leaf_cap, RtNode **out_nodes, uint32_t *out_n_nodes, int *out_root, uint32_t *out_depth, uint64_t *out_leaves,
                        float *out_sah) -> bool {
            CK(cudaMemsetAsync(d_flags, 0, sizeof(unsigned int) * (size_t)ni, st));
            CK(cudaMemsetAsync(d_leafcnt, 0, sizeof(unsigned int), st));
            k_refit_sah<<<nb, TB, 0, st>>>((int)n, sorted_vals, d_leafbox, d_kn, d_pint, d_pleaf, d_ibox, d_icost, d_coll, d_flags, leaf_cap);
            CK(cudaGetLastError());
            unsigned char root_coll = 0; float root_cost = 0.f; BoxF root_box;
            CK(cudaMemcpyAsync(&root_coll, d_coll, 1, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&root_cost, d_icost, sizeof(float), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&root_box, d_ibox, sizeof(BoxF), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            { float a = box_half_area(root_box); *out_sah = a > 0.f ? root_cost / a : 0.f; }
            if (root_coll) { *out_root = rt_leaf_ref(0, n); *out_leaves = 1; *out_n_nodes = 0; *out_depth = 0; *out_nodes = nullptr; return true; }
            RtNode *nodes = nullptr;
            CK(cudaMalloc(&nodes, sizeof(RtNode) * (size_t)ni));
            *out_nodes = nodes;                                   // owned by the mesh from here on (freed by the caller on failure)
            int init_state[4] = { 1, 0, 0, 0 };                   // frontier = { root }
            CK(cudaMemcpyAsync(d_state, init_state, sizeof init_state, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(d_front[0], 0, sizeof(int), st));
            const unsigned int fb = (ni + TB - 1) / TB;
            const int max_levels = (RT_STACK_NODES - 2) / (RT_W - 1);
            int h_state[4] = { 1, 0, 0, 0 };
            int cur = 0;
            while (h_state[0] > 0) {
                if (h_state[2] >= max_levels) { err = "BVH deeper than the traversal stack"; return false; }
                for (int b = 0; b < RT_EMIT_BATCH; b++) {
                    k_emit_count_d<<<fb, TB, 0, st>>>(d_front[cur], d_state, ni, d_kn, d_coll, d_ibox, d_cnt);
                    size_t tb = tmp_bytes;
                    CK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cnt, d_off, ni, st));
                    k_emit_write_d<<<fb, TB, 0, st>>>(d_front[cur], d_state, d_kn, d_coll, d_off, sorted_vals, d_leafbox, d_ibox, nodes,
                                                      d_front[cur ^ 1], d_leafcnt);
                    k_emit_advance<<<1, 1, 0, st>>>(d_state, d_cnt, d_off);
                    cur ^= 1;
                }
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(h_state, d_state, sizeof h_state, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
            }
            const int base = h_state[1], depth = h_state[2];
            if ((RT_W - 1) * depth + 2 > RT_STACK_NODES) { err = "BVH deeper than the traversal stack"; return false; }
            *out_n_nodes = (uint32_t)base;
            if ((size_t)base < (size_t)ni) {
                // the wide tree has far fewer nodes than the binary one it was collapsed from: shrink the allocation
                RtNode *tight = nullptr;
                CK(cudaMalloc(&tight, sizeof(RtNode) * (size_t)base));
                CK(cudaMemcpyAsync(tight, nodes, sizeof(RtNode) * (size_t)base, cudaMemcpyDeviceToDevice, st));
                CK(cudaStreamSynchronize(st));
                cudaFree(nodes);
                *out_nodes = tight;
            }
            *out_root = 0;
            *out_depth = (uint32_t)depth;
            unsigned int nl = 0;
            CK(cudaMemcpyAsync(&nl, d_leafcnt, sizeof nl, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            *out_leaves = nl;
            return true;
        };
        if (!emit(max_leaf, &m.nodes, &m.n_nodes, &m.root_ref, &m.info.max_depth, &m.info.n_leaves, &m.info.sah_cost)) return false;
        if (max_leaf_any != max_leaf) {
            float sah_any = 0.f;
            if (!emit(max_leaf_any, &m.nodes_any, &m.n_nodes_any, &m.root_ref_any, &m.info.max_depth_any, &m.info.n_leaves_any, &sah_any)) return false;
        }
    }
    if (max_leaf_any == max_leaf || n == 1) {
        m.nodes_any = m.nodes; m.n_nodes_any = m.n_nodes; m.root_ref_any = m.root_ref;
        m.info.max_depth_any = m.info.max_depth; m.info.n_leaves_any = m.info.n_leaves;
    }
    m.info.n_nodes = m.n_nodes;
    m.info.n_nodes_any = m.n_nodes_any;
#if RT_NODEQ
    if (want32) {
        // the 64-byte nodes the FP32_FAST rays fetch (rt_types.h: RtNodeQ), quantised from the float hierarchies just built
        qgrid_from_bounds(m.lo, m.hi, m.qorg, m.qstep);
        auto quantise = [&](const RtNode *src, uint32_t cnt, RtNodeQ **dst) -> bool {
            CK(cudaMalloc(dst, sizeof(RtNodeQ) * (size_t)(cnt ? cnt : 1)));
            if (cnt) k_quantise_nodes<<<(cnt + TB - 1) / TB, TB, 0, st>>>(src, cnt, *dst, m.qorg[0], m.qorg[1], m.qorg[2], m.qstep[0], m.qstep[1], m.qstep[2]);
            CK(cudaGetLastError());
            return true;
        };
        if (!quantise(m.nodes, m.n_nodes, &m.qnodes)) return false;
        if (m.nodes_any != m.nodes) { if (!quantise(m.nodes_any, m.n_nodes_any, &m.qnodes_any)) return false; }
        else m.qnodes_any = m.qnodes;
    }
#endif
    CK(cudaEventRecord(sc.ev1, st));
    CK(cudaEventSynchronize(sc.ev1));
    CK(cudaEventElapsedTime(&m.info.ms_build, sc.ev0, sc.ev1));
    sc.ok = true;
    return true;
}
