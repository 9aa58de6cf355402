// emul.cpp -- TEST-ONLY CPU replay of the device code (built by tests/, never shipped or linked into
// librt_b200.so; the product has no CPU path).  It compiles the very same RT_HD functions the sm_100a
// kernels are made of (csrc/rt_bvh.cuh builder bodies, csrc/rt_traverse.cuh traversal + shading,
// csrc/rt_exact.cuh reference-order fp64 math) with g++ and runs one loop iteration per CUDA thread.
// Purpose: catch logic errors in the kernels on the GPU-less build box, and compare the kernels'
// arithmetic against the oracle before a GPU is available.  The persistent warp-cooperative
// primary kernel (warp intrinsics) is not replayed here; it is checked on the GPU against the
// simple kernel and the oracle.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (see tests/emul_build.py)
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

// Event log (tools/simulate_warp_divergence_cpu.py): the counting hooks of the traversal code double as a recorder of the order
// in which one ray visits nodes (code 0) and leaves (code = triangles in the leaf); RT_SYNCWARP marks the start / end of a
// light's shadow query (code 254).  Off (nullptr) unless emul_shade_events is running.
static std::vector<uint8_t> *g_events = nullptr;
enum { EV_nodes = 0, EV_tris = 1 };
#define RT_COUNT(c, field, n) do { if (c) (c)->field += (n); if (g_events) g_events->push_back(EV_##field == EV_nodes ? (uint8_t)0 : (uint8_t)(n)); } while (0)
#define RT_SYNCWARP() do { if (g_events) g_events->push_back((uint8_t)254); } while (0)

#include "../../include/rt_b200.h"
#include "../../distributed-raytracer_b200/csrc/rt_types.h"
#include "../../distributed-raytracer_b200/csrc/rt_bvh.cuh"
#include "../../distributed-raytracer_b200/csrc/rt_traverse.cuh"
#include "../../distributed-raytracer_b200/csrc/rt_cull.h"
#include "host_builders_experiment.hpp"

struct EmulMesh {
    std::vector<RtNode> nodes, nodes_any;
    std::vector<RtNodeQ> qnodes, qnodes_any;
    std::vector<rt_f4> tri32;
    std::vector<double> tri64, fverts, fnormals, materials;
    std::vector<uint32_t> fmat;
    MeshView view;
    uint32_t max_depth = 0, n_leaves = 0;
};

struct EmulScene {
    std::vector<EmulMesh> meshes;
    std::vector<MeshView> views;
    std::vector<ObjRec> objs;
    uint32_t precision;
};

static void build_mesh(EmulMesh &m, int max_leaf, int max_leaf_any) {
    const uint32_t n = (uint32_t)(m.fverts.size() / 9);
    MeshView &v = m.view;
    memset(&v, 0, sizeof v);
    v.n_faces = n;
    if (n == 0) return;
    // k_mesh_bounds
    float lo[3] = { HUGE_VALF, HUGE_VALF, HUGE_VALF }, hi[3] = { -HUGE_VALF, -HUGE_VALF, -HUGE_VALF };
    for (uint32_t f = 0; f < n; f++)
        for (int q = 0; q < 3; q++)
            for (int k = 0; k < 3; k++) {
                double c = m.fverts[9 * (size_t)f + 3 * q + k];
                lo[k] = fminf(lo[k], f32_down(c));
                hi[k] = fmaxf(hi[k], f32_up(c));
            }
    BuildParams bp = make_build_params(lo, hi, v.lo, v.hi);
    std::vector<BoxF> leafbox(n);
    std::vector<unsigned long long> keys(n);
    std::vector<uint32_t> vals(n);
    for (uint32_t f = 0; f < n; f++) leaf_morton_thread(f, m.fverts.data(), bp, leafbox.data(), keys.data(), vals.data());
    // radix sort (stable) of (key, face)
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<unsigned long long> skeys(n);
    std::vector<uint32_t> svals(n);
    for (uint32_t i = 0; i < n; i++) { skeys[i] = keys[perm[i]]; svals[i] = vals[perm[i]]; }
    const char *builder = getenv("RT_B200_BUILDER");
    const bool use_sweep = builder && !strcmp(builder, "sweep"), use_sah = builder && !strcmp(builder, "sah") && n >= 2;
    std::vector<KarrasNode> kn_host(n > 1 ? n - 1 : 0);
    std::vector<int> pint_host(n > 1 ? n - 1 : 0), pleaf_host(n);
    if (use_sah) rt_sah_build_host(leafbox.data(), (int)n, svals.data(), kn_host.data(), pint_host.data(), pleaf_host.data());
    m.tri32.resize(RT_TRI32_F4 * (size_t)n);
    m.tri64.resize(10 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) gather_tri_thread(i, m.fverts.data(), svals.data(), m.tri32.data(), m.tri64.data());
    if (n == 1) {
        v.root_ref = v.root_ref_any = rt_leaf_ref(0, 1);
        m.n_leaves = 1;
    } else {
        const int ni = (int)n - 1;
        std::vector<KarrasNode> kn(ni);
        std::vector<int> pint(ni), pleaf(n);
        if (use_sah) {
            kn = kn_host; pint = pint_host; pleaf = pleaf_host;
        } else if (use_sweep) {
            std::vector<BoxF> sbox(n);
            for (uint32_t i = 0; i < n; i++) sbox[i] = leafbox[svals[i]];
            rt_sweep_build_host(sbox.data(), (int)n, kn.data(), pint.data(), pleaf.data());
        } else {
            for (int i = 0; i < ni; i++) karras_thread(i, skeys.data(), (int)n, kn.data(), pint.data(), pleaf.data());
        }
        // one wide hierarchy per leaf size over the one binary tree (bvh_build.cu: emit)
        auto emit = [&](int leaf_cap, std::vector<RtNode> &nodes, int &root_ref, uint32_t &n_nodes, uint32_t &max_depth, uint32_t &n_leaves) {
            std::vector<BoxF> ibox(ni);
            std::vector<float> icost(ni);
            std::vector<unsigned char> coll(ni);
            std::vector<unsigned int> flags(ni, 0u);
            n_leaves = 0; n_nodes = 0; max_depth = 0;
            for (int i = 0; i < (int)n; i++)
                refit_sah_thread(i, svals.data(), leafbox.data(), kn.data(), pint.data(), pleaf.data(), ibox.data(), icost.data(),
                                 coll.data(), flags.data(), leaf_cap);
            if (coll[0]) {
                root_ref = rt_leaf_ref(0, n);
                n_leaves = 1;
                return;
            }
            nodes.resize(ni);
            std::vector<int> front[2] = { std::vector<int>(ni), std::vector<int>(ni) };
            std::vector<int> cnt(ni + 1), off(ni + 1);
            front[0][0] = 0;
            int fsize = 1, base = 0, cur = 0, depth = 0;
            while (fsize > 0) {
                for (int p = 0; p < fsize; p++) emit_count_thread(p, front[cur].data(), kn.data(), coll.data(), ibox.data(), cnt.data());
                int acc = 0;
                for (int p = 0; p < fsize; p++) { off[p] = acc; acc += cnt[p]; }
                for (int p = 0; p < fsize; p++)
                    n_leaves += emit_write_thread(p, front[cur].data(), fsize, base, kn.data(), coll.data(), off.data(),
                                                  svals.data(), leafbox.data(), ibox.data(), nodes.data(), front[cur ^ 1].data());
                base += fsize;
                fsize = acc;
                cur ^= 1;
                depth++;
            }
            nodes.resize(base);
            n_nodes = (uint32_t)base;
            root_ref = 0;
            max_depth = (uint32_t)depth;
        };
        emit(max_leaf, m.nodes, v.root_ref, v.n_nodes, m.max_depth, m.n_leaves);
        if (max_leaf_any != max_leaf) {
            uint32_t d_any = 0, l_any = 0;
            emit(max_leaf_any, m.nodes_any, v.root_ref_any, v.n_nodes_any, d_any, l_any);
        }
    }
    if (max_leaf_any == max_leaf || n == 1) { v.root_ref_any = v.root_ref; v.n_nodes_any = v.n_nodes; }
    v.nodes = m.nodes.data(); v.tri32 = m.tri32.data(); v.tri64 = m.tri64.data();
    v.nodes_any = (max_leaf_any == max_leaf || n == 1) ? m.nodes.data() : m.nodes_any.data();
    // bvh_build.cu: k_quantise_nodes
    qgrid_from_bounds(v.lo, v.hi, v.qorg, v.qstep);
    m.qnodes.resize(m.nodes.size() ? m.nodes.size() : 1);
    for (uint32_t i = 0; i < m.nodes.size(); i++) quantise_node_thread(i, m.nodes.data(), m.qnodes.data(), v.qorg, v.qstep);
    v.qnodes = m.qnodes.data();
    if (v.nodes_any == v.nodes) v.qnodes_any = v.qnodes;
    else {
        m.qnodes_any.resize(m.nodes_any.size() ? m.nodes_any.size() : 1);
        for (uint32_t i = 0; i < m.nodes_any.size(); i++) quantise_node_thread(i, m.nodes_any.data(), m.qnodes_any.data(), v.qorg, v.qstep);
        v.qnodes_any = m.qnodes_any.data();
    }
}

extern "C" {

void *emul_scene_create2(const rt_scene_desc *desc, int max_leaf, int max_leaf_any);
void *emul_scene_create(const rt_scene_desc *desc, int max_leaf) { return emul_scene_create2(desc, max_leaf, max_leaf); }
void *emul_scene_create2(const rt_scene_desc *desc, int max_leaf, int max_leaf_any) {
    if (max_leaf_any < 1) max_leaf_any = max_leaf;
    EmulScene *s = new EmulScene();
    s->precision = desc->precision;
    s->meshes.resize(desc->n_meshes);
    for (uint32_t k = 0; k < desc->n_meshes; k++) {
        const rt_mesh_desc &d = desc->meshes[k];
        EmulMesh &m = s->meshes[k];
        const size_t nf = d.n_faces;
        m.fverts.resize(9 * nf);
        m.fmat.resize(nf);
        if (d.n_normals) m.fnormals.resize(9 * nf);
        for (size_t f = 0; f < nf; f++) {
            const uint32_t *fc = d.faces + 7 * f;
            for (int v = 0; v < 3; v++)
                for (int c = 0; c < 3; c++) {
                    m.fverts[9 * f + 3 * v + c] = d.verts[3 * (size_t)fc[v] + c];
                    if (d.n_normals) m.fnormals[9 * f + 3 * v + c] = d.normals[3 * (size_t)fc[3 + v] + c];
                }
            m.fmat[f] = fc[6];
        }
        m.materials.assign(d.materials, d.materials + 10 * (size_t)d.n_materials);
        build_mesh(m, max_leaf, max_leaf_any);
        for (int c = 0; c < 3; c++) { m.view.vlo[c] = HUGE_VAL; m.view.vhi[c] = -HUGE_VAL; }
        for (size_t v = 0; v < d.n_verts; v++)                   // Object.Bounds walks ALL vertices of the mesh (object.go:38-49)
            for (int c = 0; c < 3; c++) {
                m.view.vlo[c] = go_min(m.view.vlo[c], d.verts[3 * v + c]);
                m.view.vhi[c] = go_max(m.view.vhi[c], d.verts[3 * v + c]);
            }
        { const char *e = getenv("RT_B200_ANY_DEFER"); m.view.any_defer = e ? (uint32_t)std::max(1, atoi(e)) : 8u; }
        m.view.fverts = m.fverts.data();
        m.view.fnormals = d.n_normals ? m.fnormals.data() : nullptr;
        m.view.fmat = m.fmat.data();
        m.view.materials = m.materials.data();
    }
    for (auto &m : s->meshes) s->views.push_back(m.view);
    s->objs.resize(desc->n_objects);
    for (uint32_t k = 0; k < desc->n_objects; k++) {
        for (int c = 0; c < 3; c++) s->objs[k].pos[c] = desc->obj_pos[3 * (size_t)k + c];
        s->objs[k].mesh = desc->obj_mesh[k];
        s->objs[k].pad = 0;
    }
    return s;
}

void emul_scene_destroy(void *h) { delete (EmulScene *)h; }

// BVH structural facts for tests: n_nodes, max_depth, n_leaves; validates that every face is in exactly
// one leaf and that every child box contains its subtree.  Returns 0 if valid, else an error code.
int emul_bvh_check(void *h, uint32_t mesh, uint32_t *out_nodes, uint32_t *out_depth, uint32_t *out_leaves) {
    EmulScene *s = (EmulScene *)h;
    *out_nodes = 0; *out_depth = 0; *out_leaves = 0;
    if (mesh >= s->meshes.size()) return 0;          // a scene without meshes has nothing to check
    EmulMesh &m = s->meshes[mesh];
    *out_nodes = m.view.n_nodes; *out_depth = m.max_depth; *out_leaves = m.n_leaves;
    const uint32_t n = m.view.n_faces;
    if (n == 0) return 0;
    std::vector<int> seen(n, 0);
    struct Item { int ref; BoxF box; };
    std::vector<Item> stack;
    BoxF root; for (int k = 0; k < 3; k++) { root.lo[k] = m.view.lo[k]; root.hi[k] = m.view.hi[k]; }
    stack.push_back({ m.view.root_ref, root });
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        if (it.ref >= 0) {
            if ((uint32_t)it.ref >= m.view.n_nodes) return 1;
            const RtNode &nd = m.nodes[it.ref];
            int used = 0;
            for (int c = 0; c < RT_W; c++) {
                if (nd.ref[c] == RT_EMPTY) continue;
                used++;
                BoxF b = { { nd.lox[c], nd.loy[c], nd.loz[c] }, { nd.hix[c], nd.hiy[c], nd.hiz[c] } };
                for (int k = 0; k < 3; k++)
                    if (b.lo[k] < it.box.lo[k] || b.hi[k] > it.box.hi[k]) return 2;
                stack.push_back({ nd.ref[c], b });
            }
            if (used < 2) return 7;
        } else {
            uint32_t first = rt_leaf_first(it.ref), count = rt_leaf_count(it.ref);
            if (first + count > n) return 3;
            for (uint32_t k = first; k < first + count; k++) {
                int f = (int)m.tri64[10 * (size_t)k + 9];
                if (f < 0 || (uint32_t)f >= n) return 4;
                seen[f]++;
                for (int q = 0; q < 3; q++)
                    for (int c = 0; c < 3; c++) {
                        double x = m.fverts[9 * (size_t)f + 3 * q + c];
                        if (x <= (double)it.box.lo[c] || x >= (double)it.box.hi[c]) return 5;   // strictly inside (padded)
                    }
            }
        }
    }
    for (uint32_t f = 0; f < n; f++) if (seen[f] != 1) return 6;
    // quantised nodes (RtNodeQ): same refs, and every plane at least half a cell outside the float box it replaces (the decode
    // of the traversal, t = (2^23 + q) * s + c, is off by at most half a cell: rt_types.h)
    auto qcheck = [&](const std::vector<RtNode> &fn, const std::vector<RtNodeQ> &qn) -> int {
        for (size_t i = 0; i < fn.size(); i++)
            for (int c = 0; c < RT_W; c++) {
                if (qn[i].ref[c] != fn[i].ref[c]) return 8;
                if (fn[i].ref[c] == RT_EMPTY) continue;
                const float lo[3] = { fn[i].lox[c], fn[i].loy[c], fn[i].loz[c] }, hi[3] = { fn[i].hix[c], fn[i].hiy[c], fn[i].hiz[c] };
                const uint16_t ql[3] = { qn[i].x[c], qn[i].y[c], qn[i].z[c] }, qh[3] = { qn[i].x[4 + c], qn[i].y[4 + c], qn[i].z[4 + c] };
                for (int k = 0; k < 3; k++) {
                    const double o = m.view.qorg[k], st = m.view.qstep[k];
                    if (!(o + ((double)ql[k] + 0.5) * st <= (double)lo[k]) || !(o + ((double)qh[k] - 0.5) * st >= (double)hi[k])) return 9;
                    if (ql[k] == 0 || qh[k] == 65535) return 10;           // the grid has room: nothing is ever clamped
                }
            }
        return 0;
    };
    if (int rc = qcheck(m.nodes, m.qnodes)) return rc;
    if (!m.nodes_any.empty()) if (int rc = qcheck(m.nodes_any, m.qnodes_any)) return rc;
    return 0;
}

// One frame through the same pipeline the kernels implement: K2 -> K4 -> K5.
// flags: RT_FLAG_BRUTE_FORCE, RT_LAYOUT_ROWMAJOR, RT_QUANT_RGBA16_SHR8.  counters: [nodes, tris] (may be NULL).
static uint32_t *g_cost_primary = nullptr, *g_cost_shade = nullptr;   // optional per-pixel node+triangle fetch counts
static int render_core(EmulScene *s, FrameParams fp, const rt_light *lights, uint32_t flags, size_t npix, uint8_t *out_rgb,
                       double *out_colour, int32_t *out_face, int32_t *out_obj, uint32_t *out_lit, unsigned long long *counters);

static FrameParams frame_of(EmulScene *s, const rt_camera *cam, uint32_t n_lights, uint32_t W, uint32_t H) {
    FrameParams fp;
    memset(&fp, 0, sizeof fp);
    for (int k = 0; k < 3; k++) { fp.cam_pos[k] = cam->pos[k]; fp.fwd[k] = cam->forward[k]; fp.left[k] = cam->left[k]; fp.up[k] = cam->up[k]; }
    fp.tan_half = cam->tan_half_fov;
    fp.W = (int)W; fp.H = (int)H; fp.x = 0; fp.y = 0; fp.w = (int)W; fp.h = (int)H;
    fp.n_lights = (int)n_lights; fp.n_obj = (int)s->objs.size();
    return fp;
}

int emul_render(void *h, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, uint32_t W, uint32_t H, uint32_t x,
                uint32_t y, uint32_t w, uint32_t hh, uint32_t flags, uint8_t *out_rgb, double *out_colour, int32_t *out_face,
                int32_t *out_obj, uint32_t *out_lit, unsigned long long *counters) {
    EmulScene *s = (EmulScene *)h;
    FrameParams fp = frame_of(s, cam, n_lights, W, H);
    fp.x = (int)x; fp.y = (int)y; fp.w = (int)w; fp.h = (int)hh;
    return render_core(s, fp, lights, flags, (size_t)w * hh, out_rgb, out_colour, out_face, out_obj, out_lit, counters);
}

// multi-GPU tile mode of map_ray (rt_render_tiles_device): packed tiles of rank `rank`
int emul_render_tiles(void *h, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, uint32_t W, uint32_t H,
                      uint32_t tw, uint32_t th, uint32_t rank, uint32_t n_ranks, uint32_t flags, uint8_t *out_tiles) {
    EmulScene *s = (EmulScene *)h;
    FrameParams fp = frame_of(s, cam, n_lights, W, H);
    fp.tile_w = (int)tw; fp.tile_h = (int)th; fp.tiles_x = (int)((W + tw - 1) / tw);
    uint32_t n_tiles = (uint32_t)fp.tiles_x * ((H + th - 1) / th);
    fp.rank = (int)rank; fp.n_ranks = (int)n_ranks;
    fp.n_local_tiles = (int)(n_tiles > rank ? (n_tiles - rank + n_ranks - 1) / n_ranks : 0);
    return render_core(s, fp, lights, flags & ~RT_LAYOUT_ROWMAJOR, (size_t)fp.n_local_tiles * tw * th, out_tiles, nullptr, nullptr,
                       nullptr, nullptr, nullptr);
}

static int render_core(EmulScene *s, FrameParams fp, const rt_light *lights, uint32_t flags, size_t npix, uint8_t *out_rgb,
                       double *out_colour, int32_t *out_face, int32_t *out_obj, uint32_t *out_lit, unsigned long long *counters) {
    SceneView sc;
    sc.meshes = s->views.data(); sc.objs = s->objs.data(); sc.lights = reinterpret_cast<const LightRec *>(lights);
    // the replay runs one "thread" at a time: its shared-memory stack column (SceneView::trav_smem) is this array, so that the
    // hybrid shared / local stack of mesh_trace_fast (incl. the overflow into the local part) is exercised on the CPU as well
    static uint32_t s_trav[RT_SMEM_STACK_WORDS * RT_SMEM_STACK_STRIDE];
    sc.trav_smem = getenv("RT_EMUL_NO_SMEM_STACK") ? nullptr : s_trav;
    sc.tune = nullptr; sc.tune_host = nullptr; sc.defer = 0;
    const bool brute = (flags & RT_FLAG_BRUTE_FORCE) != 0, rowmajor = (flags & RT_LAYOUT_ROWMAJOR) != 0,
               sdl = (flags & RT_QUANT_RGBA16_SHR8) != 0;
    if (out_rgb) memset(out_rgb, 0, 3 * npix);
    if (out_colour) memset(out_colour, 0, sizeof(double) * 3 * npix);
    if (out_face) for (size_t k = 0; k < npix; k++) out_face[k] = -1;
    if (out_obj) for (size_t k = 0; k < npix; k++) out_obj[k] = -1;
    if (out_lit) memset(out_lit, 0, sizeof(uint32_t) * npix);
    Counters c; c.nodes = 0; c.tris = 0;
    // primary-ray culling exactly as rt_api.cu applies it (cull_params): RT_FLAG_NO_CULL / brute force switch it off
    if ((flags & (RT_FLAG_NO_CULL | RT_FLAG_BRUTE_FORCE)) == 0) rt_cull_frame(fp, s->views.data(), s->objs.data(), (int)s->objs.size());
    const uint32_t nrays = ray_count(fp);
    std::vector<HitRec> hits;
    // K2
    for (uint32_t ray = 0; ray < nrays; ray++) {
        PixMap pm = map_ray(ray, fp, false);
        if (!pm.valid) continue;
        if (fp.tile_w > 0 && block_culled(fp, pm.i - (int)(ray & 7u), pm.j - (int)((ray >> 3) & 3u))) continue;
        HitRec hr; hr.pix = ray; hr.t = 0.f;
        Counters before = c;
        struct CostNote { Counters &c; Counters b; uint32_t *dst; size_t o; ~CostNote() { if (dst) dst[o] = (uint32_t)((c.nodes - b.nodes) + (c.tris - b.tris)); } } note{ c, before, g_cost_primary, pm.out };
        if (s->precision == RT_FP64_EXACT) {
            TraceExact tr;
            dv d = primary_dir_exact(pm.i, pm.j, fp);
            if (!trace_exact(sc, fp, dv_ld(fp.cam_pos), d, brute, tr, &c)) continue;
            hr.obj = tr.obj; hr.face = tr.face;
        } else {
            // mirrors primary_dir32 of render.cu
            int hw = fp.W / 2, hh2 = fp.H / 2;
            float phw = (float)fp.tan_half, phh = phw * (float)fp.H / (float)fp.W;
            float si = phw * ((float)(hw - pm.i) - 0.5f) / (float)hw, sj = phh * ((float)(hh2 - pm.j) - 0.5f) / (float)hh2;
            float dx = (float)fp.fwd[0] + (float)fp.left[0] * si + (float)fp.up[0] * sj;
            float dy = (float)fp.fwd[1] + (float)fp.left[1] * si + (float)fp.up[1] * sj;
            float dz = (float)fp.fwd[2] + (float)fp.left[2] * si + (float)fp.up[2] * sj;
            float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            TraceFast tr;
            if (!trace_fast_nearest(sc, fp, dv_ld(fp.cam_pos), dv_make((double)(dx * inv), (double)(dy * inv), (double)(dz * inv)),
                                    true, tr, &c))
                continue;
            hr.obj = tr.obj; hr.face = tr.face; hr.t = tr.t;
        }
        hits.push_back(hr);
    }
    // K4 + K5
    const char *s32 = getenv("RT_B200_SHADE32");
    const bool shade32 = s32 && atoi(s32) != 0;        // opt-in fp32 shading of FP32_FAST (rt_traverse.cuh: shade_hit_fast32)
    for (const HitRec &hr : hits) {
        PixMap pm = map_ray(hr.pix, fp, rowmajor);
        double col[3];
        Counters before = c;
        uint32_t lit = s->precision == RT_FP64_EXACT
                           ? shade_hit<true>(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, brute, col, &c)
                           : (shade32 ? shade_hit_fast32(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, col, &c)
                                      : shade_hit<false>(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, false, col, &c));
        size_t o = pm.out;
        if (g_cost_shade) g_cost_shade[o] = (uint32_t)((c.nodes - before.nodes) + (c.tris - before.tris));
        if (out_rgb) for (int k = 0; k < 3; k++) out_rgb[3 * o + k] = quantise_exact(col[k], sdl);
        if (out_colour) for (int k = 0; k < 3; k++) out_colour[3 * o + k] = col[k];
        if (out_face) out_face[o] = hr.face;
        if (out_obj) out_obj[o] = hr.obj;
        if (out_lit) out_lit[o] = lit;
    }
    if (counters) { counters[0] = c.nodes; counters[1] = c.tris; counters[2] = hits.size(); }
    return 0;
}

// K4 event log of one frame (FP32_FAST): for every primary hit, in hit-list order (32 consecutive hits = one warp of k_shade),
// the byte sequence  254 <visits of light 0's shadow ray> 254 254 <light 1> 254 ... 255.  Returns the number of bytes
// (written up to `cap`), *n_hits = number of hits.
long long emul_shade_events(void *h, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, uint32_t W, uint32_t H,
                            uint8_t *out, unsigned long long cap, uint32_t *n_hits) {
    EmulScene *s = (EmulScene *)h;
    if (s->precision != RT_FP32_FAST) return -1;
    FrameParams fp = frame_of(s, cam, n_lights, W, H);
    fp.x = 0; fp.y = 0; fp.w = (int)W; fp.h = (int)H;
    SceneView sc;
    sc.meshes = s->views.data(); sc.objs = s->objs.data(); sc.lights = reinterpret_cast<const LightRec *>(lights);
    // the replay runs one "thread" at a time: its shared-memory stack column (SceneView::trav_smem) is this array, so that the
    // hybrid shared / local stack of mesh_trace_fast (incl. the overflow into the local part) is exercised on the CPU as well
    static uint32_t s_trav[RT_SMEM_STACK_WORDS * RT_SMEM_STACK_STRIDE];
    sc.trav_smem = getenv("RT_EMUL_NO_SMEM_STACK") ? nullptr : s_trav;
    sc.tune = nullptr; sc.tune_host = nullptr; sc.defer = 0;
    std::vector<HitRec> hits;
    Counters c; c.nodes = 0; c.tris = 0;
    const uint32_t nrays = ray_count(fp);
    const bool primary_mode = n_hits == nullptr;        // log K2 instead: every ray index (warp = 32 consecutive), sequence + 255
    std::vector<uint8_t> ev;
    unsigned long long total = 0;
    for (uint32_t ray = 0; ray < nrays; ray++) {
        PixMap pm = map_ray(ray, fp, false);
        if (!pm.valid) { if (primary_mode) { if (total < cap) out[total] = 255; total++; } continue; }
        int hw = fp.W / 2, hh2 = fp.H / 2;
        float phw = (float)fp.tan_half, phh = phw * (float)fp.H / (float)fp.W;
        float si = phw * ((float)(hw - pm.i) - 0.5f) / (float)hw, sj = phh * ((float)(hh2 - pm.j) - 0.5f) / (float)hh2;
        float dx = (float)fp.fwd[0] + (float)fp.left[0] * si + (float)fp.up[0] * sj;
        float dy = (float)fp.fwd[1] + (float)fp.left[1] * si + (float)fp.up[1] * sj;
        float dz = (float)fp.fwd[2] + (float)fp.left[2] * si + (float)fp.up[2] * sj;
        float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        TraceFast tr;
        if (primary_mode) { ev.clear(); g_events = &ev; }
        bool hit = trace_fast_nearest(sc, fp, dv_ld(fp.cam_pos), dv_make((double)(dx * inv), (double)(dy * inv), (double)(dz * inv)), true, tr, &c);
        if (primary_mode) {
            g_events = nullptr;
            ev.push_back((uint8_t)255);
            for (uint8_t b : ev) { if (total < cap) out[total] = b; total++; }
            continue;
        }
        if (!hit) continue;
        HitRec hr; hr.pix = ray; hr.obj = tr.obj; hr.face = tr.face; hr.t = tr.t;
        hits.push_back(hr);
    }
    if (primary_mode) return (long long)total;
    *n_hits = (uint32_t)hits.size();
    for (const HitRec &hr : hits) {
        PixMap pm = map_ray(hr.pix, fp, false);
        double col[3];
        ev.clear();
        g_events = &ev;
        shade_hit<false>(sc, fp, pm.i, pm.j, hr.obj, hr.face, hr.t, false, col, &c);
        g_events = nullptr;
        ev.push_back((uint8_t)255);
        for (uint8_t b : ev) { if (total < cap) out[total] = b; total++; }
    }
    return (long long)total;
}

void emul_set_cost_buffers(uint32_t *primary, uint32_t *shade) { g_cost_primary = primary; g_cost_shade = shade; }
unsigned long long emul_morton63(double x, double y, double z) { return morton63(x, y, z); }
double emul_go_max(double x, double y) { return go_max(x, y); }
double emul_go_min(double x, double y) { return go_min(x, y); }
double emul_go_pow(double x, double y) { return go_pow(x, y); }

}  // extern "C"
