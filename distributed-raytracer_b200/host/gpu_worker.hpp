// gpu_worker.hpp -- C++ mirror of the reference's worker-side interface for the render path, backed by
// the GPU C ABI.  Same names and argument meaning as (paths under /root/reference/):
//   shared/comms/comms.proto:24-45          WorkOrder{x,y,width,height,diff}, TraceResults{Colour{r,g,b}}
//   worker/distributed/main.go:27-32,46-97  Tracer{scene, screenWidth, screenHeight}, BulkTrace, Heartbeat
//   worker/distributed/main.go:101-129      register() -> NewTracer (scene arrives once; BVH built on the GPU)
//   worker/sequential/main.go:15-32         draw()
//   master/main.go:54-91                    partition()
// Error behaviour: like BulkTrace returning (nil, err) (distributed/main.go:60,74), BulkTrace returns
// false and fills `err`; the caller (master) then skips the frame (pool.go:172-174).
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "state.hpp"

namespace comms {
struct WorkOrder {                          // comms.proto:24-30
    uint32_t x = 0, y = 0, width = 0, height = 0;
    const state::EnvMutables *diff = nullptr;   // decoded gob `diff` (nullptr = GetDiff() == nil: keep the last frame state)
};
struct TraceResults {                       // comms.proto:33-41 -- packed: Results[i*height+j] = {r,g,b}
    std::vector<uint8_t> rgb;               // 3 bytes per pixel, index (i*height + j)*3  (distributed/main.go:82)
    uint32_t R(size_t k) const { return rgb[3 * k]; }
    uint32_t G(size_t k) const { return rgb[3 * k + 1]; }
    uint32_t B(size_t k) const { return rgb[3 * k + 2]; }
    size_t size() const { return rgb.size() / 3; }
};
}  // namespace comms

namespace worker {

class Tracer {                              // distributed/main.go:27-32
  public:
    // register(): the immutable scene + screen size arrive once (distributed/main.go:113-128)
    Tracer(const state::Environment &scene, uint32_t screenWidth, uint32_t screenHeight, int device, uint32_t precision);
    ~Tracer();
    Tracer(const Tracer &) = delete;
    Tracer &operator=(const Tracer &) = delete;
    bool ok() const { return handle != nullptr; }
    const std::string &error() const { return err; }
    // distributed/main.go:46-91
    bool BulkTrace(const comms::WorkOrder &req, comms::TraceResults &results, std::string &err, uint32_t flags = 0);
    // same order, results written straight into caller memory (width*height*3 bytes, e.g. a pinned buffer from
    // rt_host_alloc wrapped as a Go slice): no intermediate TraceResults copy
    bool BulkTraceInto(const comms::WorkOrder &req, uint8_t *out_rgb, std::string &err, uint32_t flags = 0);
    // same order, answered with the protobuf wire bytes of comms.TraceResults (comms.proto:34-41) encoded on the GPU: what
    // grpc-go would produce by marshalling the message BulkTrace builds (distributed/main.go:52-54,79-86).  SURVEY s8 f2.
    bool BulkTraceEncoded(const comms::WorkOrder &req, std::vector<uint8_t> &wire, std::string &err, uint32_t flags = 0);
    // distributed/main.go:94-98 (keeps the worker registered; nothing to do on the GPU side)
    void Heartbeat() {}
    rt_scene *Handle() { return handle; }
    uint32_t screenWidth, screenHeight;

  private:
    state::EnvMutables frameState(const comms::WorkOrder &req);   // per-call copy (BulkTrace is entered concurrently)
    rt_scene *handle = nullptr;
    std::mutex lastMu;
    state::EnvMutables last;                // frame state used when a WorkOrder carries no diff (guarded by lastMu)
    std::string err;
};

// worker/sequential/main.go:15-32 draw(): the whole screen into `surface` (row-major RGB8, the
// pixel (i,j) at (j*W + i)*3), background cleared to 0 first (:17).  quantiser: RT_QUANT_RGBA16_SHR8 is
// what go-sdl2's Surface.Set produces from RGBA() (colour.go:54-56); RT_QUANT_RGB255 is colour.RGB().
bool draw(Tracer &t, const state::EnvMutables &scene, std::vector<uint8_t> &surface, std::string &err,
          uint32_t quantiser = RT_QUANT_RGBA16_SHR8);

// master/main.go:54-91 partition(): recursive bisection of `area` among `workers`
// (widthKernel = heightKernel = 50, workerRedundancy = 1: master/main.go:25-31).
std::vector<comms::WorkOrder> partition(const comms::WorkOrder &area, unsigned workers, unsigned dimension, unsigned *leftover);

// SURVEY s8 row f4 -- a content-aware partition() for the master.  Same recursion as partition() (alternate axis, the same
// split of the worker count, the same 50x50 floor), but the cut is placed where the COST of the two halves is in the ratio of
// their worker counts instead of at the geometric middle.  The cost of a pixel is 1 for a primary ray that missed plus
// `hit_weight` if the pixel was hit in the PREVIOUS frame (shading + one shadow ray per light): the master already holds that
// frame (master/main.go:166-176 draws it), so no new data crosses the wire.  last_rgb: the previous frame in BulkTrace order
// (pixel (i, j) at (i*height + j)*3), nullptr = no history -> identical to partition().
std::vector<comms::WorkOrder> partitionBalanced(const comms::WorkOrder &area, unsigned workers, const uint8_t *last_rgb, uint32_t screenWidth,
                                                uint32_t screenHeight, double hit_weight);

}  // namespace worker
