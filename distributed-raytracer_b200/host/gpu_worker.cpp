// gpu_worker.cpp -- see gpu_worker.hpp
#include "gpu_worker.hpp"

namespace worker {

Tracer::Tracer(const state::Environment &scene, uint32_t w, uint32_t h, int device, uint32_t precision)
    : screenWidth(w), screenHeight(h) {
    state::FlatScene flat;
    try {
        scene.Flatten(flat, device, precision);
    } catch (const std::exception &e) {
        err = e.what();
        return;
    }
    int rc = rt_scene_create(&flat.desc, &handle);
    if (rc != RT_OK) { err = rt_last_error(nullptr); handle = nullptr; return; }
    last = scene.mutableState;
}

Tracer::~Tracer() { rt_scene_destroy(handle); }

// The frame state of ONE order, as a per-call value: the reference decodes the gob diff into a local of the RPC
// (distributed/main.go:56-64), and several BulkTrace calls run concurrently on one Tracer (a goroutine each).  An order
// without a diff renders the state of the most recent order that had one; that shared copy is only touched under the mutex.
state::EnvMutables Tracer::frameState(const comms::WorkOrder &req) {
    std::lock_guard<std::mutex> lk(lastMu);
    if (req.diff) last = *req.diff;
    return last;
}

bool Tracer::BulkTraceInto(const comms::WorkOrder &req, uint8_t *out_rgb, std::string &e, uint32_t flags) {
    if (!handle) { e = "tracer has no scene: " + err; return false; }
    if (req.width == 0 || req.height == 0) { (void)frameState(req); return true; }   // empty order: empty results, like the Go loop
    rt_camera cam;
    std::vector<rt_light> lights;
    std::vector<double> obj_pos;
    state::FrameState(frameState(req), cam, lights, obj_pos);
    int rc = rt_bulk_trace(handle, &cam, lights.data(), uint32_t(lights.size()), obj_pos.empty() ? nullptr : obj_pos.data(),
                           screenWidth, screenHeight, req.x, req.y, req.width, req.height, out_rgb,
                           (flags & ~RT_LAYOUT_ROWMAJOR) | RT_LAYOUT_COLMAJOR | RT_QUANT_RGB255);   // :79 RGB(), :82 i*height+j
    if (rc != RT_OK) { e = rt_last_error(handle); return false; }
    return true;
}

bool Tracer::BulkTrace(const comms::WorkOrder &req, comms::TraceResults &results, std::string &e, uint32_t flags) {
    results.rgb.resize(size_t(req.width) * req.height * 3);           // distributed/main.go:52-54
    return BulkTraceInto(req, results.rgb.data(), e, flags);
}

bool Tracer::BulkTraceEncoded(const comms::WorkOrder &req, std::vector<uint8_t> &wire, std::string &e, uint32_t flags) {
    wire.clear();
    if (!handle) { e = "tracer has no scene: " + err; return false; }
    if (req.width == 0 || req.height == 0) { (void)frameState(req); return true; }   // no results: the empty message
    rt_camera cam;
    std::vector<rt_light> lights;
    std::vector<double> obj_pos;
    state::FrameState(frameState(req), cam, lights, obj_pos);
    wire.resize(size_t(rt_trace_results_max_bytes(uint64_t(req.width) * req.height)));
    uint64_t len = 0;
    int rc = rt_bulk_trace_encoded(handle, &cam, lights.data(), uint32_t(lights.size()), obj_pos.empty() ? nullptr : obj_pos.data(),
                                   screenWidth, screenHeight, req.x, req.y, req.width, req.height, wire.data(), wire.size(), &len,
                                   (flags & ~RT_LAYOUT_ROWMAJOR) | RT_QUANT_RGB255);
    if (rc != RT_OK) { e = rt_last_error(handle); wire.clear(); return false; }
    wire.resize(size_t(len));
    return true;
}

bool draw(Tracer &t, const state::EnvMutables &scene, std::vector<uint8_t> &surface, std::string &e, uint32_t quantiser) {
    if (!t.ok()) { e = t.error(); return false; }
    rt_camera cam;
    std::vector<rt_light> lights;
    std::vector<double> obj_pos;
    state::FrameState(scene, cam, lights, obj_pos);
    surface.assign(size_t(t.screenWidth) * t.screenHeight * 3, 0);    // sequential/main.go:17
    int rc = rt_bulk_trace(t.Handle(), &cam, lights.data(), uint32_t(lights.size()), obj_pos.empty() ? nullptr : obj_pos.data(),
                           t.screenWidth, t.screenHeight, 0, 0, t.screenWidth, t.screenHeight, surface.data(),
                           RT_LAYOUT_ROWMAJOR | quantiser);
    if (rc != RT_OK) { e = rt_last_error(t.Handle()); return false; }
    return true;
}

static const unsigned widthKernel = 50, heightKernel = 50, workerRedundancy = 1;   // master/main.go:25-31

std::vector<comms::WorkOrder> partition(const comms::WorkOrder &area, unsigned workers, unsigned dimension, unsigned *leftover) {
    unsigned dummy = 0;
    if (!leftover) leftover = &dummy;
    if (workers / workerRedundancy < 2) {                                            // master/main.go:56-62
        *leftover = workers > workerRedundancy ? workers % workerRedundancy : 0;
        return { area };
    }
    uint32_t x = area.x, y = area.y, width = area.width, height = area.height;
    if (width <= widthKernel && height <= heightKernel) {                            // :66-68
        *leftover = workers - workerRedundancy;
        return { area };
    } else if (width <= widthKernel) dimension = 1;                                  // :69-71
    else if (height <= heightKernel) dimension = 0;                                  // :72-75
    comms::WorkOrder l = area, r = area;
    if (dimension % 2 == 0) {                                                        // :79-81
        l.x = x; l.y = y; l.width = width / 2; l.height = height;
        r.x = x + width / 2; r.y = y; r.width = width / 2 + width % 2; r.height = height;
    } else {                                                                         // :82-85
        l.x = x; l.y = y; l.width = width; l.height = height / 2;
        r.x = x; r.y = y + height / 2; r.width = width; r.height = height / 2 + height % 2;
    }
    unsigned remainder = 0;
    std::vector<comms::WorkOrder> left = partition(l, workers / 2 + workers % 2, (dimension + 1) % 2, &remainder);   // :88
    std::vector<comms::WorkOrder> right = partition(r, workers / 2 + remainder, (dimension + 1) % 2, &remainder);     // :89
    left.insert(left.end(), right.begin(), right.end());
    *leftover = remainder;
    return left;
}

}  // namespace worker
