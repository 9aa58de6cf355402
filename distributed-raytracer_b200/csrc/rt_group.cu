// rt_group.cu -- one worker process driving SEVERAL GPUs of the box through the C ABI (include/rt_b200.h, rt_group_*).
//
// The reference dispatches ONE BulkTrace order per worker (master/pool/pool.go:148-197 -> worker/distributed/main.go:46-91).
// A worker on an 8-GPU box is one Go process: it must not have to re-implement a multi-process launcher to use the other seven
// GPUs.  rt_group_create registers the scene on every device (what register() does once, distributed/main.go:101-129), and
// rt_group_bulk_trace answers one order with all of them:
//
//   * the region is cut into stripes that are contiguous in the reply's layout (BulkTrace's column-major layout: 8 pixel
//     columns x the region's height; row-major: 4 pixel rows x its width), stripe t belongs to device t mod N -- the busy
//     middle of the frame is spread evenly (a partition() rectangle per GPU is not: DESIGN.md s5);
//   * every device renders its stripes in place with the fused frame kernel and moves them with ONE strided copy-engine
//     transfer over ITS OWN PCIe link straight into the caller's buffer (pinned: rt_host_alloc) or into the call's pinned
//     staging frame;
//   * no collective, no inter-GPU traffic, no device-side hand-off: the calling thread enqueues N short command sequences and
//     waits for N streams.  Calls are thread-safe and concurrent (one group lane each), like rt_bulk_trace.
//
// This is the orchestration that round 1 had in Python over torch.distributed (parallel.py, mode "host"); parallel.py's
// one-process-per-GPU form stays for the torchrun benchmark, where the driver launches one rank per GPU.
#include "rt_internal.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

namespace {

struct GroupLane {
    bool busy = false;
    std::vector<cudaStream_t> streams;     // one per device
    std::vector<uint8_t *> d_local;        // per device: the region's framebuffer in its final layout
    std::vector<size_t> cap_local;
    uint8_t *h_stage = nullptr;            // pinned staging frame (caller buffers that are not pinned)
    size_t cap_stage = 0;
};

thread_local std::string g_group_err;

// One enqueue thread per device: the command sequence of an order (frame state, clear, frame kernel, strided copy) costs
// ~20 us of host time per device; issued by the calling thread alone that is 8 x 20 us at 8 GPUs, more than the 0.12 ms the
// GPUs then need.  The workers only ENQUEUE (asynchronous); the caller waits for the streams.  A worker spins briefly after its
// last task before it sleeps, so a worker that is being fed stays hot and one that is not costs nothing.
struct DevWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::atomic<int> pending{ 0 };
    bool stop = false;

    void post(std::function<void()> f) {
        { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(f)); pending.fetch_add(1, std::memory_order_release); }
        cv.notify_one();
    }
    void run() {
        for (;;) {
            std::function<void()> f;
            // spin ~200 us for the next task before sleeping
            const auto t0 = std::chrono::steady_clock::now();
            while (pending.load(std::memory_order_acquire) == 0) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
                std::this_thread::yield();
            }
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;                      // stop requested and nothing left
                f = std::move(q.front());
                q.pop_front();
                pending.fetch_sub(1, std::memory_order_relaxed);
            }
            f();
        }
    }
};

}  // namespace

struct rt_group {
    std::mutex mu;
    std::condition_variable cv;
    std::string err;
    std::vector<int> devices;
    std::vector<rt_scene *> scenes;
    std::vector<double> obj_pos;           // registered object positions (an order without positions renders these)
    std::vector<std::unique_ptr<GroupLane>> lanes;
    std::vector<std::unique_ptr<DevWorker>> workers;   // one per device entry except the first (the caller enqueues that one itself)
    uint32_t stripe_cols = 8, stripe_rows = 4;
};

namespace {

int gfail(rt_group *g, int code, const std::string &msg) {
    if (g) { std::lock_guard<std::mutex> lk(g->mu); g->err = msg; }
    g_group_err = msg;
    return code;
}

void destroy_group(rt_group *g) {
    for (auto &wp : g->workers) {
        { std::lock_guard<std::mutex> lk(wp->mu); wp->stop = true; }
        wp->cv.notify_all();
        if (wp->th.joinable()) wp->th.join();
    }
    g->workers.clear();
    for (auto &lp : g->lanes) {
        for (size_t d = 0; d < g->devices.size(); d++) {
            cudaSetDevice(g->devices[d]);
            if (d < lp->streams.size() && lp->streams[d]) { cudaStreamSynchronize(lp->streams[d]); cudaStreamDestroy(lp->streams[d]); }
            if (d < lp->d_local.size()) cudaFree(lp->d_local[d]);
        }
        if (lp->h_stage) cudaFreeHost(lp->h_stage);
    }
    g->lanes.clear();
    for (rt_scene *s : g->scenes) rt_scene_destroy(s);
    g->scenes.clear();
}

GroupLane *acquire(rt_group *g, int *index) {
    std::unique_lock<std::mutex> lk(g->mu);
    for (;;) {
        for (size_t k = 0; k < g->lanes.size(); k++)
            if (!g->lanes[k]->busy) { g->lanes[k]->busy = true; *index = (int)k; return g->lanes[k].get(); }
        g->cv.wait(lk);
    }
}
void release(rt_group *g, GroupLane *L) {
    { std::lock_guard<std::mutex> lk(g->mu); L->busy = false; }
    g->cv.notify_one();
}

}  // namespace

extern "C" {

int rt_group_create(const rt_scene_desc *desc, const int32_t *devices, uint32_t n_devices, rt_group **out) {
    g_group_err.clear();
    if (!desc || !out) return gfail(nullptr, RT_ERR_INVALID, "rt_group_create: NULL argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return gfail(nullptr, RT_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    }
    std::vector<int> devs;
    if (!devices || n_devices == 0) {                       // every GPU of the box
        for (int d = 0; d < ndev; d++) devs.push_back(d);
    } else {
        for (uint32_t k = 0; k < n_devices; k++) {
            if (devices[k] < 0 || devices[k] >= ndev) return gfail(nullptr, RT_ERR_INVALID, "rt_group_create: device ordinal out of range");
            devs.push_back(devices[k]);
        }
    }
    if (devs.size() > 32) return gfail(nullptr, RT_ERR_INVALID, "rt_group_create: at most 32 devices");
    rt_group *g = new (std::nothrow) rt_group();
    if (!g) return gfail(nullptr, RT_ERR_OOM, "out of host memory");
    g->devices = devs;
    g->obj_pos.assign(desc->obj_pos, desc->obj_pos + 3 * (size_t)desc->n_objects);
    int n_lanes = 8;
    if (const char *v = getenv("RT_B200_GROUP_LANES")) if (*v) n_lanes = atoi(v);
    if (n_lanes < 1) n_lanes = 1;
    if (n_lanes > 32) n_lanes = 32;
    // every device registers the scene: its own upload + LBVH build (the builds are deterministic, so all devices hold the same tree)
    for (int d : devs) {
        rt_scene_desc dd = *desc;
        dd.device = d;
        rt_scene *s = nullptr;
        // the per-device scenes need lane 0 + one lane per group lane
        char buf[16];
        const char *old = getenv("RT_B200_LANES");
        std::string keep = old ? old : "";
        snprintf(buf, sizeof buf, "%d", n_lanes + 1);
        setenv("RT_B200_LANES", buf, 1);
        int rc = rt_scene_create(&dd, &s);
        if (old) setenv("RT_B200_LANES", keep.c_str(), 1); else unsetenv("RT_B200_LANES");
        if (rc != RT_OK) {
            std::string msg = std::string("rt_group_create: device ") + std::to_string(d) + ": " + rt_last_error(nullptr);
            destroy_group(g);
            delete g;
            return gfail(nullptr, rc, msg);
        }
        g->scenes.push_back(s);
    }
    for (int k = 0; k < n_lanes; k++) {
        g->lanes.emplace_back(new GroupLane());
        GroupLane &L = *g->lanes.back();
        L.streams.assign(devs.size(), nullptr);
        L.d_local.assign(devs.size(), nullptr);
        L.cap_local.assign(devs.size(), 0);
        for (size_t d = 0; d < devs.size(); d++) {
            cudaError_t e = cudaSetDevice(devs[d]);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&L.streams[d], cudaStreamNonBlocking);
            if (e != cudaSuccess) {
                std::string msg = std::string("rt_group_create: stream: ") + cudaGetErrorString(e);
                destroy_group(g);
                delete g;
                return gfail(nullptr, RT_ERR_CUDA, msg);
            }
        }
    }
    for (size_t d = 1; d < devs.size(); d++) {
        g->workers.emplace_back(new DevWorker());
        DevWorker *w = g->workers.back().get();
        w->th = std::thread([w] { w->run(); });
    }
    *out = g;
    return RT_OK;
}

void rt_group_destroy(rt_group *g) {
    if (!g) return;
    {   // wait for the calls in flight
        std::unique_lock<std::mutex> lk(g->mu);
        for (;;) {
            bool any = false;
            for (auto &lp : g->lanes) any = any || lp->busy;
            if (!any) break;
            g->cv.wait(lk);
        }
    }
    destroy_group(g);
    delete g;
}

int rt_group_device_count(const rt_group *g) { return g ? (int)g->devices.size() : 0; }

rt_scene *rt_group_scene(rt_group *g, uint32_t k) { return (g && k < g->scenes.size()) ? g->scenes[k] : nullptr; }

const char *rt_group_last_error(const rt_group *g) {
    if (!g || !g_group_err.empty()) return g_group_err.c_str();
    static thread_local std::string ret;
    { std::lock_guard<std::mutex> lk(const_cast<rt_group *>(g)->mu); ret = g->err; }
    return ret.c_str();
}

int rt_group_bulk_trace(rt_group *g, const rt_camera *cam, const rt_light *lights, uint32_t n_lights, const double *obj_pos, uint32_t W,
                        uint32_t H, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint8_t *out_rgb, uint32_t flags) {
    g_group_err.clear();
    if (!g) return gfail(nullptr, RT_ERR_INVALID, "NULL group");
    if (!out_rgb || !cam) return gfail(g, RT_ERR_INVALID, "rt_group_bulk_trace: NULL argument");
    if (W == 0 || H == 0 || w == 0 || h == 0 || (uint64_t)x + w > W || (uint64_t)y + h > H)
        return gfail(g, RT_ERR_INVALID, "empty region or region exceeds the screen");
    if (flags >> 24) return gfail(g, RT_ERR_INVALID, "rt_group_bulk_trace: RT_LANE does not apply (the group picks the lanes)");
    const uint32_t n = (uint32_t)g->devices.size();
    const bool rowmajor = (flags & RT_LAYOUT_ROWMAJOR) != 0;
    const uint32_t stripe = rowmajor ? g->stripe_rows : g->stripe_cols;
    const size_t bytes = (size_t)w * h * 3;
    int idx = 0;
    GroupLane *L = acquire(g, &idx);
    int rc = RT_OK;
    std::string msg;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == RT_OK) { rc = e == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA; msg = std::string(what) + ": " + cudaGetErrorString(e); }
        return e == cudaSuccess;
    };
    // destination of the copy-engine transfers: the caller's buffer if it is pinned, else this lane's pinned staging frame
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, out_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    uint8_t *dst = out_rgb;
    if (!pinned) {
        if (bytes > L->cap_stage) {
            if (L->h_stage) cudaFreeHost(L->h_stage);
            L->h_stage = nullptr; L->cap_stage = 0;
            if (cu(cudaHostAlloc((void **)&L->h_stage, bytes, cudaHostAllocPortable), "cudaHostAlloc")) L->cap_stage = bytes;
        }
        dst = L->h_stage;
    }
    const double *pos = obj_pos ? obj_pos : (g->obj_pos.empty() ? nullptr : g->obj_pos.data());
    // per-device enqueue: device 0 by this thread, the others by their workers, all at once
    std::vector<int> d_rc(n, RT_OK);
    std::vector<std::string> d_msg(n);
    std::vector<char> d_launched(n, 0);
    std::atomic<uint32_t> done{ 0 };
    auto enqueue = [&](uint32_t d) {
        cudaError_t e = cudaSetDevice(g->devices[d]);
        if (e == cudaSuccess && bytes > L->cap_local[d]) {
            cudaFree(L->d_local[d]);
            L->d_local[d] = nullptr; L->cap_local[d] = 0;
            e = cudaMalloc((void **)&L->d_local[d], bytes);
            if (e == cudaSuccess) L->cap_local[d] = bytes;
        }
        if (e != cudaSuccess) {
            d_rc[d] = e == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA;
            d_msg[d] = std::string("device ") + std::to_string(g->devices[d]) + ": " + cudaGetErrorString(e);
        } else {
            int r = rt_internal_render_stripes(g->scenes[d], cam, lights, n_lights, pos, W, H, x, y, w, h, stripe, d, n, L->d_local[d], dst, nullptr, 0,
                                               L->streams[d], (flags & 0x00ffffffu) | RT_LANE((uint32_t)idx + 1u));
            if (r != RT_OK) { d_rc[d] = r; d_msg[d] = std::string("device ") + std::to_string(g->devices[d]) + ": " + rt_last_error(g->scenes[d]); }
            else d_launched[d] = 1;
        }
        done.fetch_add(1, std::memory_order_release);
    };
    if (rc == RT_OK) {
        for (uint32_t d = 1; d < n; d++) g->workers[d - 1]->post([&enqueue, d] { enqueue(d); });
        enqueue(0);
        while (done.load(std::memory_order_acquire) < n) std::this_thread::yield();      // ~20 us of enqueue work per device
        for (uint32_t d = 0; d < n; d++)
            if (d_rc[d] != RT_OK && rc == RT_OK) { rc = d_rc[d]; msg = d_msg[d]; }
    }
    for (uint32_t d = 0; d < n; d++) {                      // every stream that got work is drained, error or not
        if (!d_launched[d]) continue;
        cudaSetDevice(g->devices[d]);
        cu(cudaStreamSynchronize(L->streams[d]), "cudaStreamSynchronize");
    }
    if (rc == RT_OK && !pinned) memcpy(out_rgb, dst, bytes);
    release(g, L);
    return rc == RT_OK ? RT_OK : gfail(g, rc, msg);
}

}  // extern "C"
