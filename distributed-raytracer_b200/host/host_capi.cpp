// host_capi.cpp -- C exports of the C++ host mirror so that the Python tests / bench can drive the
// reference-facing interface (EnvironmentFromFile, Tracer.BulkTrace, partition) through ctypes.
#include <algorithm>
#include <cstring>

#include "gpu_worker.hpp"

namespace {
thread_local std::string g_err;
struct EnvBox { state::Environment env; state::FlatScene flat; };
}  // namespace

extern "C" {

const char *rth_last_error(void) { return g_err.c_str(); }

void *rth_environment_from_file(const char *path) {
    try {
        EnvBox *b = new EnvBox();
        b->env = state::EnvironmentFromFile(path);
        b->env.Flatten(b->flat, 0, RT_FP32_FAST);
        return b;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void rth_environment_free(void *h) { delete (EnvBox *)h; }

// flat views for tests (pointers stay valid until rth_environment_free)
const rt_scene_desc *rth_environment_desc(void *h) { return &((EnvBox *)h)->flat.desc; }
uint32_t rth_environment_n_lights(void *h) { return uint32_t(((EnvBox *)h)->env.mutableState.Lights.size()); }
// fills cam (rt_camera), lights (n*6 doubles) and obj_pos (n_obj*3 doubles)
void rth_environment_frame(void *h, rt_camera *cam, double *lights, double *obj_pos) {
    std::vector<rt_light> l;
    std::vector<double> p;
    state::FrameState(((EnvBox *)h)->env.mutableState, *cam, l, p);
    if (lights) std::memcpy(lights, l.data(), sizeof(rt_light) * l.size());
    if (obj_pos) std::memcpy(obj_pos, p.data(), sizeof(double) * p.size());
}
// camera motion (camera.go:62-154) on the environment's mutable state
void rth_camera_move(void *h, double distance, int f, int b, int l, int r, int u, int d) { ((EnvBox *)h)->env.mutableState.Cam.Move(distance, f, b, l, r, u, d); }
void rth_camera_yaw(void *h, double theta) { ((EnvBox *)h)->env.mutableState.Cam.Yaw(theta); }
void rth_camera_pitch(void *h, double theta) { ((EnvBox *)h)->env.mutableState.Cam.Pitch(theta); }
void rth_camera_set(void *h, const double *pos, const double *dir, double fov) {
    try { ((EnvBox *)h)->env.mutableState.Cam = state::NewCamera({ pos[0], pos[1], pos[2] }, { dir[0], dir[1], dir[2] }, fov); }
    catch (const std::exception &e) { g_err = e.what(); }
}

void *rth_tracer_new(void *h, uint32_t screen_w, uint32_t screen_h, int device, uint32_t precision) {
    worker::Tracer *t = new worker::Tracer(((EnvBox *)h)->env, screen_w, screen_h, device, precision);
    if (!t->ok()) { g_err = t->error(); delete t; return nullptr; }
    return t;
}
void rth_tracer_free(void *t) { delete (worker::Tracer *)t; }
rt_scene *rth_tracer_handle(void *t) { return ((worker::Tracer *)t)->Handle(); }

// Tracer.BulkTrace with the environment's current mutable state as the order's diff; out: w*h*3 bytes, (i*h+j)*3
int rth_bulk_trace(void *t, void *h, uint32_t x, uint32_t y, uint32_t w, uint32_t hgt, uint8_t *out, uint32_t flags) {
    comms::WorkOrder req;
    req.x = x; req.y = y; req.width = w; req.height = hgt;
    req.diff = h ? &((EnvBox *)h)->env.mutableState : nullptr;
    std::string err;
    if (!((worker::Tracer *)t)->BulkTraceInto(req, out, err, flags)) { g_err = err; return -1; }
    return 0;
}
// Tracer.BulkTraceEncoded: the order answered as comms.TraceResults wire bytes; out must hold 11*w*h bytes; returns the length or -1
long long rth_bulk_trace_encoded(void *t, void *h, uint32_t x, uint32_t y, uint32_t w, uint32_t hgt, uint8_t *out, uint32_t flags) {
    comms::WorkOrder req;
    req.x = x; req.y = y; req.width = w; req.height = hgt;
    req.diff = h ? &((EnvBox *)h)->env.mutableState : nullptr;
    std::string err;
    std::vector<uint8_t> wire;
    if (!((worker::Tracer *)t)->BulkTraceEncoded(req, wire, err, flags)) { g_err = err; return -1; }
    std::memcpy(out, wire.data(), wire.size());
    return (long long)wire.size();
}
// draw(): whole screen, row-major
int rth_draw(void *t, void *h, uint8_t *surface, uint32_t quantiser) {
    std::vector<uint8_t> s;
    std::string err;
    if (!worker::draw(*(worker::Tracer *)t, ((EnvBox *)h)->env.mutableState, s, err, quantiser)) { g_err = err; return -1; }
    std::memcpy(surface, s.data(), s.size());
    return 0;
}
// partition(): out = n*4 uint32 (x,y,w,h); returns n (<= cap)
uint32_t rth_partition(uint32_t W, uint32_t H, uint32_t workers, uint32_t *out, uint32_t cap) {
    comms::WorkOrder area;
    area.width = W; area.height = H;
    unsigned leftover = 0;
    std::vector<comms::WorkOrder> v = worker::partition(area, workers, 0, &leftover);
    uint32_t n = 0;
    for (const comms::WorkOrder &o : v) {
        if (n >= cap) break;
        out[4 * n] = o.x; out[4 * n + 1] = o.y; out[4 * n + 2] = o.width; out[4 * n + 3] = o.height;
        n++;
    }
    return n;
}

// register a GPU scene straight from the scene file (every model goes to the device as OBJ text and is parsed there).
// Returns the scene handle (rt_scene_destroy frees it) or nullptr; cam / lights / obj_pos as rth_environment_frame fills them
// (lights: up to max_lights * 6 doubles, *n_lights receives the count; obj_pos: up to max_objs * 3 doubles).
rt_scene *rth_gpu_scene_from_file(const char *scene_path, int device, uint32_t precision, rt_camera *cam, double *lights, uint32_t max_lights,
                                  uint32_t *n_lights, double *obj_pos, uint32_t max_objs, uint32_t *n_objs) {
    rt_scene *s = nullptr;
    state::EnvMutables em;
    std::string err;
    if (state::GpuSceneFromFile(scene_path, device, precision, &s, em, err) != RT_OK) { g_err = err; return nullptr; }
    std::vector<rt_light> l;
    std::vector<double> p;
    state::FrameState(em, *cam, l, p);
    *n_lights = uint32_t(l.size()); *n_objs = uint32_t(p.size() / 3);
    if (lights) std::memcpy(lights, l.data(), sizeof(rt_light) * std::min<size_t>(l.size(), max_lights));
    if (obj_pos) std::memcpy(obj_pos, p.data(), sizeof(double) * 3 * std::min<size_t>(p.size() / 3, max_objs));
    return s;
}
// the host half alone (file read + mtllib / usemtl scan + MTL parse), for timing: returns the bytes read, or -1
long long rth_obj_text_probe(const char *obj_path, uint32_t *n_materials, uint32_t *n_usemtl) {
    try {
        state::ObjTextMesh t = state::ObjTextFromFile(obj_path);
        *n_materials = uint32_t(t.materials.size() / 10); *n_usemtl = uint32_t(t.usemtl_off.size());
        return (long long)t.text.size();
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// partitionBalanced(): like rth_partition, cut by the cost of the previous frame (last_rgb: W*H*3 bytes in BulkTrace order)
uint32_t rth_partition_balanced(uint32_t W, uint32_t H, uint32_t workers, const uint8_t *last_rgb, double hit_weight, uint32_t *out, uint32_t cap) {
    comms::WorkOrder area;
    area.width = W; area.height = H;
    std::vector<comms::WorkOrder> v = worker::partitionBalanced(area, workers, last_rgb, W, H, hit_weight);
    uint32_t n = 0;
    for (const comms::WorkOrder &o : v) {
        if (n >= cap) break;
        out[4 * n] = o.x; out[4 * n + 1] = o.y; out[4 * n + 2] = o.width; out[4 * n + 3] = o.height;
        n++;
    }
    return n;
}

}  // extern "C"
