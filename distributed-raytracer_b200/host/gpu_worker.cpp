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

namespace {

// summed-area table of the per-pixel cost (column-major like the frame: index i*(H+1) + j over (W+1) x (H+1))
struct CostTable {
    uint32_t W = 0, H = 0;
    std::vector<double> sat;
    double sum(uint32_t x, uint32_t y, uint32_t w, uint32_t h) const {
        auto at = [&](uint32_t i, uint32_t j) { return sat[size_t(i) * (H + 1) + j]; };
        return at(x + w, y + h) - at(x, y + h) - at(x + w, y) + at(x, y);
    }
};

void balanced(const CostTable &ct, const comms::WorkOrder &area, unsigned workers, unsigned dimension, unsigned *leftover,
              std::vector<comms::WorkOrder> &out) {
    if (workers / workerRedundancy < 2) {                                            // master/main.go:56-62
        *leftover = workers > workerRedundancy ? workers % workerRedundancy : 0;
        out.push_back(area);
        return;
    }
    const uint32_t x = area.x, y = area.y, width = area.width, height = area.height;
    if (width <= widthKernel && height <= heightKernel) {                            // :66-68
        *leftover = workers - workerRedundancy;
        out.push_back(area);
        return;
    } else if (width <= widthKernel) dimension = 1;
    else if (height <= heightKernel) dimension = 0;
    const unsigned wl = workers / 2 + workers % 2;                                   // workers of the left half (:88)
    const double want = ct.sum(x, y, width, height) * double(wl) / double(workers);
    const bool vertical = dimension % 2 == 0;
    const uint32_t extent = vertical ? width : height;
    // smallest cut whose left part carries at least its share of the cost (binary search on the monotone prefix cost),
    // kept at least one pixel away from both ends
    uint32_t lo = 1, hi = extent - 1;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        const double c = vertical ? ct.sum(x, y, mid, height) : ct.sum(x, y, width, mid);
        if (c < want) lo = mid + 1; else hi = mid;
    }
    const uint32_t cut = lo;
    comms::WorkOrder l = area, r = area;
    if (vertical) { l.width = cut; r.x = x + cut; r.width = width - cut; }
    else { l.height = cut; r.y = y + cut; r.height = height - cut; }
    unsigned remainder = 0;
    balanced(ct, l, wl, (dimension + 1) % 2, &remainder, out);
    balanced(ct, r, workers / 2 + remainder, (dimension + 1) % 2, &remainder, out);
    *leftover = remainder;
}

}  // namespace

std::vector<comms::WorkOrder> partitionBalanced(const comms::WorkOrder &area, unsigned workers, const uint8_t *last_rgb, uint32_t W,
                                                uint32_t H, double hit_weight) {
    if (!last_rgb || W == 0 || H == 0) return partition(area, workers, 0, nullptr);
    CostTable ct;
    ct.W = W; ct.H = H;
    ct.sat.assign(size_t(W + 1) * (H + 1), 0.0);
    for (uint32_t i = 0; i < W; i++) {
        double col = 0.0;                                        // running cost of pixel column i
        for (uint32_t j = 0; j < H; j++) {
            const uint8_t *p = last_rgb + (size_t(i) * H + j) * 3;
            col += 1.0 + ((p[0] | p[1] | p[2]) ? hit_weight : 0.0);
            ct.sat[size_t(i + 1) * (H + 1) + (j + 1)] = ct.sat[size_t(i) * (H + 1) + (j + 1)] + col;
        }
    }
    std::vector<comms::WorkOrder> out;
    unsigned leftover = 0;
    balanced(ct, area, workers, 0, &leftover, out);
    return out;
}

}  // namespace worker
