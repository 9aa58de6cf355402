// obj_ingest.cu -- SURVEY s8 row f3: Wavefront OBJ text -> indexed mesh arrays ON THE GPU.
//
// The reference loads a model with gwob (a serial tokenizer) and then walks every triangle corner through three value-keyed Go
// maps to de-duplicate vertices, normals and materials (/root/reference/shared/state/mesh.go:109-213, :146-203) before it inserts
// the faces one by one into an R-tree (:208).  None of that is needed to render: de-duplication only saves memory, and the
// faces go into an LBVH that is built on the device anyway.  Here the OBJ file's bytes go to the GPU as they are and
//
//   k_obj_classify   one thread per line: kind (v / vn / f / other), corners of a face line
//   cub scans        running counts of v / vn lines and of emitted triangles = output index of every line
//   k_obj_parse      one thread per line: decimal -> float32 -> float64 (gwob parses with 32-bit precision and the reference
//                    widens, mesh.go:176 Coord64), face corners "v", "v/t", "v//n", "v/t/n" with 1-based or negative
//                    (relative) indices, quad -> (0,1,2) + (2,3,0), the face's material from the usemtl table
//   k_obj_normalise  vn -> unit length in the reference's operation order (vector.go:50-53; mesh.go:202)
//
// produce exactly the arrays rt_mesh_desc describes (face index = emission order), already on the device,
// where rt_flatten_faces + the LBVH build take over.  Decimal parsing is exact for what OBJ files contain (<= 15 significant
// digits, |exponent| <= 22: one correctly rounded double operation); a value whose double lands exactly on a float32 rounding
// boundary -- where rounding the double could differ from rounding the decimal -- and anything outside that range is put on a
// short fix-up list that the host resolves with strtof, so the result is bit-identical to a correctly rounded float32 parse.
#include "rt_internal.h"
#include "rt_exact.cuh"
#include <cub/cub.cuh>
#include <cstdlib>

namespace {

enum : uint32_t { OBJ_OTHER = 0, OBJ_V = 1, OBJ_VN = 2, OBJ_F = 3 };
enum : uint32_t { OBJ_ERR_FACE_ARITY = 1, OBJ_ERR_INDEX = 2, OBJ_ERR_SYNTAX = 3, OBJ_ERR_MIXED_NORMALS = 4, OBJ_ERR_FIXUPS = 5 };
constexpr uint32_t OBJ_MAX_FIXUPS = 65536;

struct ObjFix { uint32_t array; uint32_t index; uint64_t offset; };     // array 0 = verts, 1 = normals; index = element; offset = token start

__device__ __forceinline__ bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r'; }
__device__ __forceinline__ bool is_eol(char c) { return c == '\n' || c == '\0'; }

__constant__ double c_pow10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18,
                                     1e19, 1e20, 1e21, 1e22 };

// Parses one decimal number at p (< end).  Returns the position after it (p itself if there is no number), the value rounded to
// float32 and widened, and `exact` = the float32 rounding is provably the correctly rounded one.
__device__ const char *parse_f32(const char *p, const char *end, double &out, bool &exact) {
    const char *s = p;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
    unsigned long long mant = 0;
    int digits = 0, dropped = 0, frac = 0;
    bool any = false, dot = false;
    for (; s < end; s++) {
        const char c = *s;
        if (c >= '0' && c <= '9') {
            any = true;
            if (mant == 0 && c == '0') { if (dot) frac++; continue; }         // leading zeros carry no significance
            if (digits < 19) { mant = mant * 10ull + (unsigned long long)(c - '0'); digits++; if (dot) frac++; }
            else { dropped++; if (!dot) frac--; }                               // beyond 19 digits: scale only
        } else if (c == '.' && !dot) dot = true;
        else break;
    }
    if (!any) { out = 0.0; exact = true; return p; }
    int e10 = -frac;
    if (s < end && (*s == 'e' || *s == 'E')) {
        const char *t = s + 1;
        bool eneg = false;
        if (t < end && (*t == '-' || *t == '+')) { eneg = *t == '-'; t++; }
        int ev = 0;
        bool edig = false;
        for (; t < end && *t >= '0' && *t <= '9'; t++) { edig = true; if (ev < 10000) ev = ev * 10 + (*t - '0'); }
        if (edig) { e10 += eneg ? -ev : ev; s = t; }
    }
    exact = dropped == 0 && mant < (1ull << 53) && e10 >= -22 && e10 <= 22;
    double d = (double)mant;
    if (mant != 0) {
        if (e10 >= -22 && e10 <= 22) d = e10 < 0 ? __ddiv_rn(d, c_pow10[-e10]) : __dmul_rn(d, c_pow10[e10]);
        else d = d * pow(10.0, (double)e10);
    }
    const float f = (float)d;                                                   // round to nearest even
    // a double exactly half way between two floats (or in the float subnormal / overflow range): let the host decide
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    if ((bits & 0x1fffffffull) == 0x10000000ull) exact = false;
    if (mant != 0 && (fabs(d) < 1.2e-38 || fabs(d) > 3.4e38)) exact = false;
    out = neg ? -(double)f : (double)f;
    return s;
}

__device__ const char *skip_ws(const char *p, const char *end) {
    while (p < end && is_ws(*p)) p++;
    return p;
}
__device__ const char *parse_int(const char *p, const char *end, long long &v, bool &ok) {
    const char *s = p;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
    long long x = 0;
    bool any = false;
    for (; s < end && *s >= '0' && *s <= '9'; s++) { any = true; if (x < (1ll << 40)) x = x * 10 + (*s - '0'); }
    ok = any;
    v = neg ? -x : x;
    return any ? s : p;
}

__global__ void k_obj_mark(const char *__restrict__ text, uint64_t n, uint8_t *__restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || text[i - 1] == '\n') ? 1 : 0;
}

// counts[3*l + {0,1,2}] = this line's number of v lines, vn lines, emitted triangles
__global__ void k_obj_classify(const char *__restrict__ text, uint64_t n, const uint64_t *__restrict__ line_start, uint32_t n_lines,
                               uint32_t *__restrict__ cnt_v, uint32_t *__restrict__ cnt_vn, uint32_t *__restrict__ cnt_tri, unsigned int *err) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const char *end = text + n;
    const char *p = skip_ws(text + line_start[l], end);
    uint32_t v = 0, vn = 0, tri = 0;
    if (p + 1 < end && p[0] == 'v' && is_ws(p[1])) v = 1;
    else if (p + 2 < end && p[0] == 'v' && p[1] == 'n' && is_ws(p[2])) vn = 1;
    else if (p + 1 < end && p[0] == 'f' && is_ws(p[1])) {
        uint32_t corners = 0;
        p++;
        for (;;) {
            p = skip_ws(p, end);
            if (p >= end || is_eol(*p)) break;
            corners++;
            while (p < end && !is_ws(*p) && !is_eol(*p)) p++;
        }
        if (corners == 3) tri = 1;
        else if (corners == 4) tri = 2;
        else atomicMax(err, OBJ_ERR_FACE_ARITY);           // gwob: triangles and quads only
    }
    cnt_v[l] = v; cnt_vn[l] = vn; cnt_tri[l] = tri;
}

__global__ void k_obj_parse(const char *__restrict__ text, uint64_t n, const uint64_t *__restrict__ line_start, uint32_t n_lines,
                            const uint32_t *__restrict__ cnt_v, const uint32_t *__restrict__ cnt_vn, const uint32_t *__restrict__ cnt_tri,
                            const uint32_t *__restrict__ off_v, const uint32_t *__restrict__ off_vn, const uint32_t *__restrict__ off_tri,
                            const uint64_t *__restrict__ usemtl_off, const uint32_t *__restrict__ usemtl_mat, uint32_t n_usemtl,
                            uint32_t default_mat, uint32_t total_v, uint32_t total_vn, double *__restrict__ verts, double *__restrict__ normals,
                            uint32_t *__restrict__ faces, unsigned int *err, unsigned int *has_normals, unsigned int *no_normals,
                            ObjFix *fix, unsigned int *n_fix) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const char *end = text + n;
    const uint64_t ls = line_start[l];
    const char *p = skip_ws(text + ls, end);
    if (cnt_v[l] || cnt_vn[l]) {
        const bool isn = cnt_vn[l] != 0;
        double *dst = isn ? normals + 3 * (size_t)off_vn[l] : verts + 3 * (size_t)off_v[l];
        p += isn ? 2 : 1;
        for (int c = 0; c < 3; c++) {
            p = skip_ws(p, end);
            double val;
            bool exact;
            const char *q = parse_f32(p, end, val, exact);
            if (q == p) { atomicMax(err, OBJ_ERR_SYNTAX); val = 0.0; exact = true; }
            if (!exact) {
                const unsigned int k = atomicAdd(n_fix, 1u);
                if (k < OBJ_MAX_FIXUPS) { fix[k].array = isn ? 1u : 0u; fix[k].index = 3u * (isn ? off_vn[l] : off_v[l]) + (uint32_t)c; fix[k].offset = (uint64_t)(p - text); }
            }
            dst[c] = val;
            p = q;
        }
        return;
    }
    if (!cnt_tri[l]) return;
    // face line: the material in force is that of the last usemtl line before it
    uint32_t mat = default_mat;
    {
        uint32_t lo = 0, hi = n_usemtl;                      // first entry with offset > ls
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (usemtl_off[mid] <= ls) lo = mid + 1; else hi = mid; }
        if (lo > 0) mat = usemtl_mat[lo - 1];
    }
    uint32_t cv[4] = { 0, 0, 0, 0 }, cn[4] = { 0, 0, 0, 0 };
    const uint32_t corners = cnt_tri[l] + 2;
    const long long nv_before = off_v[l], nvn_before = off_vn[l];          // v / vn lines above this one (relative indices)
    bool any_n = false, all_n = true;
    p++;
    for (uint32_t c = 0; c < corners; c++) {
        p = skip_ws(p, end);
        long long vi = 0, ni = 0;
        bool ok = false, okn = false;
        p = parse_int(p, end, vi, ok);
        if (!ok) atomicMax(err, OBJ_ERR_SYNTAX);
        if (p < end && *p == '/') {
            p++;
            long long ti; bool okt;
            p = parse_int(p, end, ti, okt);                 // texture index: ignored
            if (p < end && *p == '/') { p++; p = parse_int(p, end, ni, okn); }
        }
        while (p < end && !is_ws(*p) && !is_eol(*p)) p++;
        const long long v0 = vi > 0 ? vi - 1 : nv_before + vi;
        if (v0 < 0 || v0 >= (long long)total_v) { atomicMax(err, OBJ_ERR_INDEX); cv[c] = 0; } else cv[c] = (uint32_t)v0;
        if (okn) {
            const long long n0 = ni > 0 ? ni - 1 : nvn_before + ni;
            if (n0 < 0 || n0 >= (long long)total_vn) { atomicMax(err, OBJ_ERR_INDEX); cn[c] = 0; } else cn[c] = (uint32_t)n0;
            any_n = true;
        } else all_n = false;
    }
    if (any_n) atomicOr(has_normals, 1u);
    if (!all_n) atomicOr(no_normals, 1u);
    uint32_t *f = faces + 7 * (size_t)off_tri[l];
    f[0] = cv[0]; f[1] = cv[1]; f[2] = cv[2]; f[3] = cn[0]; f[4] = cn[1]; f[5] = cn[2]; f[6] = mat;
    if (corners == 4) {                                      // gwob's quad split: (v0 v1 v2) + (v2 v3 v0)
        f += 7;
        f[0] = cv[2]; f[1] = cv[3]; f[2] = cv[0]; f[3] = cn[2]; f[4] = cn[3]; f[5] = cn[0]; f[6] = mat;
    }
}

// mesh.go:202: vertexNormals are stored normalised (vector.go:50-53, reference operation order)
__global__ void k_obj_normalise(double *__restrict__ normals, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dv v = dv_norm(dv_make(normals[3 * (size_t)i], normals[3 * (size_t)i + 1], normals[3 * (size_t)i + 2]));
    normals[3 * (size_t)i] = v.x; normals[3 * (size_t)i + 1] = v.y; normals[3 * (size_t)i + 2] = v.z;
}

struct Tmp {
    std::vector<void *> ptrs;
    template <class T> cudaError_t alloc(T **p, size_t bytes) {
        cudaError_t e = cudaMalloc((void **)p, bytes ? bytes : 16);
        if (e == cudaSuccess) ptrs.push_back((void *)*p);
        return e;
    }
    ~Tmp() { for (void *p : ptrs) cudaFree(p); }
};

}  // namespace

#define CKO(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// text: HOST pointer to the OBJ file's bytes.  On success *d_verts / *d_normals / *d_faces are device arrays (caller frees) with
// n_verts / n_normals (0 = the mesh has no usable normals: flat-normal path) / n_faces entries; ms = device time of the ingest.
bool rt_obj_ingest(const char *text, uint64_t n, const uint64_t *usemtl_off, const uint32_t *usemtl_mat, uint32_t n_usemtl, uint32_t default_mat,
                   cudaStream_t st, double **d_verts, uint32_t *n_verts, double **d_normals, uint32_t *n_normals, uint32_t **d_faces,
                   uint32_t *n_faces, float *ms, std::string &err) {
    *d_verts = nullptr; *d_normals = nullptr; *d_faces = nullptr; *n_verts = *n_normals = *n_faces = 0; *ms = 0.f;
    if (n == 0) return true;
    if (n >= (1ull << 40)) { err = "OBJ text too large"; return false; }
    Tmp tmp;
    char *d_text = nullptr; uint8_t *d_flag = nullptr; uint64_t *d_lines = nullptr; unsigned int *d_n_lines = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CKO(cudaEventCreate(&e0)); CKO(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ e0, e1 };
    CKO(tmp.alloc(&d_text, n + 1));
    CKO(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, st));
    CKO(cudaMemsetAsync(d_text + n, 0, 1, st));
    CKO(cudaEventRecord(e0, st));
    CKO(tmp.alloc(&d_flag, n));
    CKO(tmp.alloc(&d_lines, sizeof(uint64_t) * (n / 2 + 2)));           // a line is at least 1 byte + '\n' ... except empty lines
    CKO(tmp.alloc(&d_n_lines, sizeof(unsigned int) * 2));
    k_obj_mark<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_text, n, d_flag);
    CKO(cudaGetLastError());
    // the line-start list may need up to n entries in a file of empty lines: count first, then select
    void *d_cub = nullptr; size_t cub_bytes = 0;
    unsigned long long *d_cnt64 = nullptr;
    CKO(tmp.alloc(&d_cnt64, 16));
    {
        size_t b1 = 0, b2 = 0;
        CKO(cub::DeviceReduce::Sum(nullptr, b1, d_flag, d_cnt64, (int64_t)n, st));
        cub::CountingInputIterator<uint64_t> it(0);
        CKO(cub::DeviceSelect::Flagged(nullptr, b2, it, d_flag, d_lines, d_cnt64 + 1, (int64_t)n, st));
        cub_bytes = b1 > b2 ? b1 : b2;
    }
    CKO(tmp.alloc(&d_cub, cub_bytes));
    {
        size_t b = cub_bytes;
        CKO(cub::DeviceReduce::Sum(d_cub, b, d_flag, d_cnt64, (int64_t)n, st));
    }
    unsigned long long h_lines = 0;
    CKO(cudaMemcpyAsync(&h_lines, d_cnt64, sizeof h_lines, cudaMemcpyDeviceToHost, st));
    CKO(cudaStreamSynchronize(st));
    if (h_lines > 0xfffffff0ull) { err = "OBJ file has too many lines"; return false; }
    const uint32_t n_lines = (uint32_t)h_lines;
    if (h_lines > n / 2 + 2) {                                           // rare: re-allocate the list at its true size
        CKO(tmp.alloc(&d_lines, sizeof(uint64_t) * (size_t)h_lines));
    }
    {
        size_t b = cub_bytes;
        cub::CountingInputIterator<uint64_t> it(0);
        CKO(cub::DeviceSelect::Flagged(d_cub, b, it, d_flag, d_lines, d_cnt64 + 1, (int64_t)n, st));
    }
    uint32_t *d_cnt = nullptr, *d_off = nullptr; unsigned int *d_err = nullptr;
    CKO(tmp.alloc(&d_cnt, sizeof(uint32_t) * 3 * ((size_t)n_lines + 1)));
    CKO(tmp.alloc(&d_off, sizeof(uint32_t) * 3 * ((size_t)n_lines + 1)));
    CKO(tmp.alloc(&d_err, sizeof(unsigned int) * 8));
    CKO(cudaMemsetAsync(d_err, 0, sizeof(unsigned int) * 8, st));
    CKO(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * 3 * ((size_t)n_lines + 1), st));
    uint32_t *cnt_v = d_cnt, *cnt_vn = d_cnt + (n_lines + 1), *cnt_tri = d_cnt + 2 * (size_t)(n_lines + 1);
    uint32_t *off_v = d_off, *off_vn = d_off + (n_lines + 1), *off_tri = d_off + 2 * (size_t)(n_lines + 1);
    const unsigned lb = (n_lines + 255) / 256;
    k_obj_classify<<<lb, 256, 0, st>>>(d_text, n, d_lines, n_lines, cnt_v, cnt_vn, cnt_tri, d_err);
    CKO(cudaGetLastError());
    {
        size_t b = 0;
        CKO(cub::DeviceScan::ExclusiveSum(nullptr, b, cnt_v, off_v, (int)(n_lines + 1), st));
        void *d_scan = nullptr;
        CKO(tmp.alloc(&d_scan, b));
        for (int k = 0; k < 3; k++) {                        // the extra (zero) element at the end receives the total
            size_t bb = b;
            CKO(cub::DeviceScan::ExclusiveSum(d_scan, bb, d_cnt + k * (size_t)(n_lines + 1), d_off + k * (size_t)(n_lines + 1), (int)(n_lines + 1), st));
        }
    }
    uint32_t tot[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) CKO(cudaMemcpyAsync(&tot[k], d_off + k * (size_t)(n_lines + 1) + n_lines, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    unsigned int h_err[8] = { 0 };
    CKO(cudaMemcpyAsync(h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost, st));
    CKO(cudaStreamSynchronize(st));
    if (h_err[0] == OBJ_ERR_FACE_ARITY) { err = "OBJ face with unsupported vertex count (gwob: triangles and quads only)"; return false; }
    const uint32_t nv = tot[0], nvn = tot[1], ntri = tot[2];
    // outputs
    double *verts = nullptr, *normals = nullptr; uint32_t *faces = nullptr;
    uint64_t *d_uoff = nullptr; uint32_t *d_umat = nullptr; ObjFix *d_fix = nullptr;
    auto drop = [&]() { cudaFree(verts); cudaFree(normals); cudaFree(faces); };
#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); drop(); return false; } } while (0)
    CKD(cudaMalloc((void **)&verts, sizeof(double) * 3 * (size_t)(nv ? nv : 1)));
    CKD(cudaMalloc((void **)&normals, sizeof(double) * 3 * (size_t)(nvn ? nvn : 1)));
    CKD(cudaMalloc((void **)&faces, sizeof(uint32_t) * 7 * (size_t)(ntri ? ntri : 1)));
    CKD(tmp.alloc(&d_uoff, sizeof(uint64_t) * (n_usemtl ? n_usemtl : 1)));
    CKD(tmp.alloc(&d_umat, sizeof(uint32_t) * (n_usemtl ? n_usemtl : 1)));
    CKD(tmp.alloc(&d_fix, sizeof(ObjFix) * OBJ_MAX_FIXUPS));
    if (n_usemtl) {
        CKD(cudaMemcpyAsync(d_uoff, usemtl_off, sizeof(uint64_t) * n_usemtl, cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d_umat, usemtl_mat, sizeof(uint32_t) * n_usemtl, cudaMemcpyHostToDevice, st));
    }
    k_obj_parse<<<lb, 256, 0, st>>>(d_text, n, d_lines, n_lines, cnt_v, cnt_vn, cnt_tri, off_v, off_vn, off_tri, d_uoff, d_umat, n_usemtl, default_mat,
                                    nv, nvn, verts, normals, faces, d_err, d_err + 1, d_err + 2, d_fix, d_err + 3);
    CKD(cudaGetLastError());
    CKD(cudaMemcpyAsync(h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));
    if (h_err[0] == OBJ_ERR_INDEX) { err = "OBJ vertex / normal index out of range"; drop(); return false; }
    if (h_err[0] == OBJ_ERR_SYNTAX) { err = "OBJ syntax error in a v / vn / f line"; drop(); return false; }
    if (h_err[1] && h_err[2]) { err = "OBJ mixes faces with and without normal indices (not supported by the GPU ingest)"; drop(); return false; }
    // fix-up list: values whose float32 rounding the device could not prove (see parse_f32): the host parses them with strtof
    const unsigned int nfix = h_err[3];
    if (nfix > OBJ_MAX_FIXUPS) { err = "OBJ: too many numbers outside the exact device parsing range"; drop(); return false; }
    if (nfix) {
        std::vector<ObjFix> hf(nfix);
        CKD(cudaMemcpyAsync(hf.data(), d_fix, sizeof(ObjFix) * nfix, cudaMemcpyDeviceToHost, st));
        CKD(cudaStreamSynchronize(st));
        for (const ObjFix &f : hf) {
            // the token ends at whitespace; strtof stops there by itself (the text is NUL- or newline-terminated by the caller's buffer)
            char buf[64];
            size_t k = 0;
            while (k < sizeof buf - 1 && f.offset + k < n && !(text[f.offset + k] == ' ' || text[f.offset + k] == '\t' || text[f.offset + k] == '\r' || text[f.offset + k] == '\n')) { buf[k] = text[f.offset + k]; k++; }
            buf[k] = 0;
            const double val = (double)strtof(buf, nullptr);
            CKD(cudaMemcpyAsync((f.array ? normals : verts) + f.index, &val, sizeof(double), cudaMemcpyHostToDevice, st));
            CKD(cudaStreamSynchronize(st));
        }
    }
    const bool has_n = h_err[1] != 0 && nvn > 0;
    if (has_n) {
        k_obj_normalise<<<(nvn + 255) / 256, 256, 0, st>>>(normals, nvn);
        CKD(cudaGetLastError());
    }
    CKD(cudaEventRecord(e1, st));
    CKD(cudaEventSynchronize(e1));
    CKD(cudaEventElapsedTime(ms, e0, e1));
#undef CKD
    *d_verts = verts; *n_verts = nv;
    *d_normals = normals; *n_normals = has_n ? nvn : 0;
    *d_faces = faces; *n_faces = ntri;
    return true;
}
