#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json):
Mrays/s (primary + shadow rays) on the 4K render of Suzanne subdivided to 991 232 triangles, 2 lights, FP32_FAST.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one frame = per-frame state upload (camera, lights) + K2 primary + K4 shadow/shade + K5 framebuffer
(+ at N>1: gather of the interleaved tile buffers to rank 0 and the de-interleave kernel).
`value`  : whole-job Mrays/s with the scene and BVH resident in HBM and the framebuffer left on the device.
`e2e`    : the same frames through the reference-facing call (Tracer.BulkTrace -> rt_bulk_trace) with HOST buffers:
           frame state H2D and the w*h*3-byte result D2H (pinned) inside the timed region.
`roofline`, `cpu_baseline`: see DESIGN.md "Measurement".
Rays per frame = W*H primary + (hit pixels) * (lights) shadow (tracer.go:60-64), identical for GPU and CPU arms.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs that are bench workloads: C3 is the configuration the metric is quoted on (the default); C4 is the
# BVH-build / traversal stress (--config c4), the one scene whose working set (~2 GB) is HBM-sized.
CONFIGS = {
    "c3": "C3: Suzanne subdivided x5 (991232 tris) @3840x2160, 2 lights, example camera, FP32_FAST",
    "c4": "C4: 10000000-triangle random soup (seed 0x5EED, flat normals) @3840x2160, 8 white lights on a ring, example camera, FP32_FAST",
}
WORKLOAD = CONFIGS["c3"]
CONFIG = "c3"
W, H = 3840, 2160
SUBDIV = 5
N_LIGHTS = 2
C4_TRIS = 10_000_000
METRIC = "Mrays/s (primary+shadow)"
REPLICAS = 8      # frames in flight (lanes of the one resident scene, or scene copies with --replicated-scenes); 12 at N > 1
TILE = (32, 16)
E2E_TILE = (8, 16)    # host-mode e2e: stripes of 8 pixel columns x full height (contiguous in the column-major layout)
SCENE_JSON = os.path.join(ROOT, "tests", "golden", "example", "scene.json")


def build_workload(config=None):
    """Scene arrays of the bench configuration, produced through the product's own loader (C++ host mirror) + workloads."""
    from distributed_raytracer_b200 import host, workloads
    env = host.Environment(SCENE_JSON)
    sc = env.scene_arrays()
    if (config or CONFIG) == "c4":
        sc["meshes"] = [workloads.triangle_soup(C4_TRIS)]
        sc["lights"] = workloads.ring_lights(8)
    else:
        sc["meshes"] = [workloads.subdivide(sc["meshes"][0], SUBDIV)]
        sc["lights"] = sc["lights"][:N_LIGHTS]
    env.close()
    return sc


def build_workload_reference_arm(config):
    """The same scene arrays WITHOUT touching the product: the oracle's own loader (oracle/scene_loader.py) and the numpy
    generators of workloads.py loaded by file path (importing the package would not load a native library either, but the
    reference arm keeps clear of it altogether)."""
    import importlib.util
    from oracle import scene_loader as sl
    spec = importlib.util.spec_from_file_location("rt_workloads", os.path.join(ROOT, "distributed-raytracer_b200", "workloads.py"))
    workloads = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(workloads)
    sc = sl.environment_from_file(SCENE_JSON)
    if config == "c4":
        sc["meshes"] = [workloads.triangle_soup(C4_TRIS)]
        sc["lights"] = workloads.ring_lights(8)
    else:
        sc["meshes"] = [workloads.subdivide(sc["meshes"][0], SUBDIV)]
        sc["lights"] = sc["lights"][:N_LIGHTS]
    if "tan_half_fov" not in sc:
        from oracle import oracle as orc
        sc["tan_half_fov"] = orc.go_tan(float(sc["camera"][12]) / 2.0)
    return sc


def config_dict(**kw):
    """The `config` object of the JSON line: the SAME keys on the repo's arm and on the reference arm (values that do not apply
    to an arm are null)."""
    keys = ("workload", "rays_per_frame", "hit_pixels", "frames_per_s", "l2", "scene_copies", "scene_create_ms", "frames_in_flight",
            "host_enqueue_ms_per_step", "precision", "parallelism", "bvh", "render_kernel", "note")
    d = {k: None for k in keys}
    d.update(kw)
    assert set(d) == set(keys), sorted(set(d) ^ set(keys))
    return d


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  Polls NVML (the source nvidia-smi reads) every few ms from a
    thread, because a timed region of a few hundred ms is too short for `nvidia-smi -lms`; falls back to nvidia-smi."""

    def __init__(self, gpu_index=0):
        self.gpu_index, self.rows, self.stop_flag, self.thread, self.max_mhz = gpu_index, [], False, None, None

    def _poll_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop_flag:
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0, time.perf_counter()))
            except Exception:
                pass
            time.sleep(0.004)
        nv.nvmlShutdown()

    def start(self):
        try:
            import pynvml  # noqa: F401
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[3] <= t1]
            # an NVML query can take longer than a short timed region: fall back to the samples nearest to it
            self.rows = inside if len(inside) >= 2 else sorted(self.rows, key=lambda r: abs(r[3] - 0.5 * (t0 + t1)))[:3]
        if not self.rows:
            return self._nvidia_smi_once()
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names.items() if bits & b), "samples": len(sm),
                "power_w_max": max(r[2] for r in self.rows), "source": "nvml poll (4 ms period) over the timed region"}

    def _nvidia_smi_once(self):
        try:
            out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                           "--format=csv,noheader,nounits", "-i", str(self.gpu_index)], text=True)
            c = [x.strip() for x in out.strip().split(",")]
            return {"sm_mhz": float(c[0]), "sm_max_mhz": float(c[1]), "reasons": [c[2]], "samples": 1, "source": "nvidia-smi (single sample)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def oracle_cpu_rate(sc, n_threads, budget_s, mode_name="rtree"):
    """Times the CPU oracle (reference-shaped R-tree walk, /root/reference/worker/sequential semantics) on a bounded
    sample of the same workload (whole pixel columns).  Returns (Mrays/s, rays, seconds, stride)."""
    from oracle import oracle as orc
    S = orc.Scene(sc)
    S.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc.get("tan_half_fov", float("nan")))
    mode = orc.RTREE if mode_name == "rtree" else orc.CULLED
    if mode == orc.RTREE:
        S.build_rtrees()                                  # scene load (mesh.go:208), not part of the frame time
    t0 = time.perf_counter()
    _, _, _, st = S.render_region(W, H, mode=mode, n_threads=n_threads, step=512, phase=5, want_colour=False, want_face=False)
    probe = time.perf_counter() - t0
    rate = (st["primary_rays"] + st["shadow_rays"]) / max(probe, 1e-6)
    full_rays = W * H * 1.2
    stride = max(1, int(round(full_rays / max(rate * budget_s, 1.0))))
    # the sample: every stride-th band of 16 whole pixel columns (the reference loop walks a column at a time,
    # distributed/main.go:67-69, so the CPU keeps its cache locality; the bands spread over the whole screen keep the
    # hit/miss mix of the full frame)
    band, rays, dt = 16, 0, 0.0
    for x0 in range(0, W - band + 1, band * stride):
        t0 = time.perf_counter()
        _, _, _, st = S.render_region(W, H, x=x0, y=0, w=band, h=H, mode=mode, n_threads=n_threads, want_colour=False, want_face=False)
        dt += time.perf_counter() - t0
        rays += st["primary_rays"] + st["shadow_rays"]
    S.close()
    return rays / dt / 1e6, rays, dt, stride


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The Go binary cannot be built in this
    image (no Go toolchain), so this is the oracle port, reference-shaped (R-tree (3,2,5) candidate walk with
    box.go's predicate, no tmax, full nearest-hit shadow queries), run on all host threads."""
    if rank != 0:
        return
    # only the oracle is built and loaded here: no CUDA library of the repo is mapped into this process
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], env=dict(os.environ, CC="gcc", CXX="g++"))
    sc = build_workload_reference_arm(args.config)
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    S = orc.Scene(sc)
    S.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    S.build_rtrees()
    # Size the per-step sample so that the whole run ends in ~100 s whatever K is.  A step renders a band of whole pixel
    # COLUMNS (x0 .. x0+band, full height), the bands cycling over the screen: that is the order the reference's own loop
    # walks (i outer, j inner: distributed/main.go:67-69), so the CPU keeps the cache locality it has in the real program
    # (a strided pixel sample does not: measured 1.3 instead of 2.5 Mrays/s on 8 cores), and K steps together cover the
    # frame uniformly.
    t0 = time.perf_counter()
    _, _, _, st = S.render_region(W, H, mode=orc.RTREE, n_threads=0, step=512, phase=5, want_colour=False, want_face=False)
    rate = (st["primary_rays"] + st["shadow_rays"]) / max(time.perf_counter() - t0, 1e-6)
    per_step_s = max(0.05, min(6.0, 100.0 / max(1, args.steps + args.warmup)))
    band = int(min(W, max(64, round(W * rate * per_step_s / (W * H * 1.2)))))
    n_bands = (W + band - 1) // band
    times, rays = [], 0
    for it in range(args.warmup + args.steps):
        x0 = min((it % n_bands) * band, W - band)
        t0 = time.perf_counter()
        _, _, _, st = S.render_region(W, H, x=x0, y=0, w=band, h=H, mode=orc.RTREE, n_threads=0, want_colour=False, want_face=False)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            rays += st["primary_rays"] + st["shadow_rays"]
    total = sum(times)
    value = rays / total / 1e6
    sample = ("a band of %d pixel columns x %d rows of the 4K frame per step (%d bands cycling over the screen, the reference "
              "loop's own order), R-tree walk" % (band, H, n_bands))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(workload=WORKLOAD, precision="float64 (the reference's arithmetic)",
                              parallelism="%d host threads, each a band of pixel columns (one reference worker per thread)" % cores,
                              note="C restatement of worker/sequential's Go path (oracle port), not the Go binary"),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    global REPLICAS, E2E_TILE, WORKLOAD, CONFIG
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="bench workload (default: c3, the configuration the metric is quoted on)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=0, help="frames in flight (default %d at N = 1, 12 at N > 1)" % REPLICAS)
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e loop (experiments only)")
    ap.add_argument("--replicated-scenes", action="store_true",
                    help="every frame in flight renders from its OWN resident copy of the scene (a frame's inputs were last touched "
                         "that many frames earlier); default: one registered scene, one lane per frame in flight")
    ap.add_argument("--mode", default="peer", choices=["peer", "peer_nccl", "gather"],
                    help="N>1: peer-memory tile push with device-side hand-off flags | same push with an NCCL barrier | NCCL gather")
    ap.add_argument("--e2e-mode", default="host", choices=["host", "peer"],
                    help="N>1 e2e: every GPU writes its tiles into a host-shared pinned framebuffer | NVLink push + D2H from rank 0")
    ap.add_argument("--e2e-tile", default="", help="tile WxH of the host-mode e2e renderer (default %dx%d)" % E2E_TILE)
    ap.add_argument("--persistent", action="store_true", help="use the persistent warp-cooperative primary kernel (RT_FLAG_PERSISTENT)")
    ap.add_argument("--e2e-callers", type=int, default=4, help="N = 1: concurrent rt_bulk_trace callers of the e2e loop")
    ap.add_argument("--pipeline", action="store_true", help="three kernels per frame (K2 -> K4 -> K5, RT_FLAG_PIPELINE) instead of the fused frame kernel")
    args = ap.parse_args()
    CONFIG, WORKLOAD = args.config, CONFIGS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # frames in flight: 8 saturate one GPU; across GPUs a few more absorb the skew between the ranks (measured at N = 8:
    # 0.1183 ms/frame with 8, 0.1156 with 12; profiles/r01_n8_sweep.jsonl)
    REPLICAS = args.replicas if args.replicas > 0 else (REPLICAS if world == 1 else 12)
    args.shared_scene = not args.replicated_scenes
    if args.e2e_tile:
        E2E_TILE = tuple(int(v) for v in args.e2e_tile.lower().split("x"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    import __graft_entry__ as ge

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # device-side hand-off waits give up after this (default 2 s, the reference's order time-out): the bench box is shared, a
        # rank that is descheduled for a few seconds must not cost the line (a region hit by a time-out is re-measured anyway)
        os.environ.setdefault("RT_B200_SYNC_TIMEOUT_MS", "8000")
        dist.init_process_group("nccl", device_id=dev)
    # rank 0 (re)builds the native libraries if needed; the other ranks must not load them while they are being written
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import host, parallel

    render_flags = rt.FLAG_PERSISTENT if args.persistent else (rt.FLAG_PIPELINE if args.pipeline else 0)
    sc = build_workload() if rank == 0 else None
    if world > 1:
        # scene broadcast once over NCCL/NVLink; every rank builds the same LBVH locally
        frame = [(sc["camera"], sc["lights"], sc["tan_half_fov"])] if rank == 0 else [None]
        dist.broadcast_object_list(frame, src=0)
        cam13, lights, tan_half = frame[0]
        sc_r = parallel.broadcast_scene(sc, src=0, device=dev)
    else:
        cam13, lights, tan_half = sc["camera"], sc["lights"], sc["tan_half_fov"]
        sc_r = sc

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.Stream(dev)            # the "main" stream: flushes, timing events, serial passes
    torch.cuda.set_stream(stream)

    # The render pipeline: REPLICAS frames in flight, one stream each; frame k goes to lane k % REPLICAS of the one resident
    # scene (--replicated-scenes: to its own copy of the scene).
    r = parallel.TileShardedRenderer(sc_r, W, H, precision=rt.FP32_FAST, tile=TILE, device_index=local_rank, replicas=REPLICAS,
                                     layout_flags=rt.LAYOUT_COLMAJOR, mode=args.mode, shared_scene=args.shared_scene)
    torch.cuda.synchronize(dev)
    # register() cost, measured on a SECOND handle (the first one also pays CUDA module loading): indexed mesh H2D +
    # device-side flatten + LBVH build + lane set-up.  The GPU idles at ~120 MHz and DVFS needs a few hundred ms of work to reach
    # its load clocks: without this ramp the same build reads anything between 7 and 490 ms (round 1: 21.7 ms in BENCH, 58 ms in SCALE)
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 0.4:
        flush.zero_()
        torch.cuda.synchronize(dev)
    t_create = time.perf_counter()
    g2 = rt.GpuScene(sc_r, device=local_rank, precision=rt.FP32_FAST)
    torch.cuda.synchronize(dev)
    t_create = time.perf_counter() - t_create
    g2_bvh = g2.bvh_info()
    g2.close()
    g = r.backend.scenes[0]
    bvh = g2_bvh                                 # the build timed at load clocks (the first handle was built on a cold GPU)
    # kernels of the repo per frame: the fused frame kernel (--pipeline: K2 K4 K5) (+ push-tiles (+ frame-complete on rank 0) |
    # untile on rank 0); rank 0's count is the one reported
    render_kernels = 3 if args.pipeline else 1
    launches_per_step = render_kernels if world == 1 else render_kernels + ((2 if rank == 0 else 1) if args.mode == "peer" else 1)

    frame_state = rt.FrameState(cam13, lights, tan_half_fov=tan_half)   # camera + lights marshalled once; handed over per frame

    def step():
        return r.render_frame(frame_state, flags=render_flags)

    # ---- counting pass (outside the timed region, serial): algorithmic traversal work of one frame on this rank ----
    g.set_frame(cam13, lights, tan_half_fov=tan_half)
    scratch = r.frame[0] if world == 1 else torch.empty(rt.tiles_per_rank(W, H, TILE[0], TILE[1], world) * TILE[0] * TILE[1] * 3, dtype=torch.uint8, device=dev)

    def serial_frame(flags):
        if world == 1:
            g.render_region_device(W, H, 0, 0, W, H, scratch.data_ptr(), None, flags)
        else:
            g.render_tiles_device(W, H, TILE[0], TILE[1], rank, world, scratch.data_ptr(), None, flags)
    serial_frame(render_flags | rt.FLAG_COUNT)
    cnt = g.stats()
    hits_local, prim_local = cnt["hit_pixels"], cnt["primary_rays"]
    if world > 1:
        t = torch.tensor([float(hits_local)], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        hits_frame = int(t[0].item())
    else:
        hits_frame = hits_local
    rays_frame = W * H + hits_frame * len(lights)         # tile padding beyond the screen is not a ray

    # ---- per-kernel durations (serial frames, L2 flushed before each, CUDA events on the launching stream) ----
    # both launch shapes are timed: the frame kernel that runs by default and the three-kernel pipeline it replaced
    kern_ms = {"frame": [], "primary": [], "shade": [], "write": []}
    fused_default = not (args.pipeline or args.persistent)
    for _ in range(3 + min(args.steps, 20)):
        flush.zero_()
        torch.cuda.synchronize(dev)
        serial_frame(rt.FLAG_PIPELINE | (rt.FLAG_PERSISTENT if args.persistent else 0))
        st = g.stats()
        for name in ("primary", "shade", "write"):
            kern_ms[name].append(st["ms_" + name])
        flush.zero_()
        torch.cuda.synchronize(dev)
        serial_frame(rt.FLAG_FUSED)
        kern_ms["frame"].append(g.stats()["ms_total"])
    kern_ms = {k: v[3:] for k, v in kern_ms.items()}

    # ---- N > 1: the sharded frame must BE the single-GPU frame, byte for byte (both the device-resident path of `value` and the
    # host path of `e2e`); checked before anything is timed, the run aborts otherwise ----
    frame_identical = None
    ref_frame = None
    if world > 1:
        if rank == 0:
            ref_frame = torch.empty(W * H * 3, dtype=torch.uint8, device=dev)
            g.render_region_device(W, H, 0, 0, W, H, ref_frame.data_ptr(), None, render_flags | rt.LAYOUT_COLMAJOR)
            torch.cuda.synchronize(dev)
        f = r.render_frame(frame_state, flags=render_flags)
        r.wait()
        torch.cuda.synchronize(dev)
        dist.barrier()
        ok = torch.tensor([1 if (rank != 0 or torch.equal(f, ref_frame)) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        frame_identical = {"value_path_" + args.mode: bool(ok.item())}
        if not ok.item():
            raise SystemExit("bench: the %d-GPU frame differs from the single-GPU frame (mode %s)" % (world, args.mode))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- clock ramp (untimed, not a step): the GPU idles at ~120 MHz; give DVFS ~0.4 s of work to reach its load clocks ----
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 0.4:
        flush.zero_()
        torch.cuda.synchronize(dev)
    # ---- lane priming (untimed, not a step): the first frame on a lane allocates its scratch buffers (cudaMalloc waits for
    # the device), so every frame slot is used once before anything is measured; then the W warm-up steps ----
    if world > 1:
        dist.barrier()                                    # the ranks enter the frame pipeline together (a rank > 2 s late is a hand-off time-out)
    for _ in range(REPLICAS):
        step()
    r.wait()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    r.check_sync(" (lane priming)")
    for _ in range(args.warmup):
        step()
    r.wait()
    torch.cuda.synchronize(dev)
    r.check_sync(" (warm-up)")

    # ---- timed region: exactly K steps, barrier + synchronize on both sides, device timing, max over ranks ----
    # A device-side hand-off wait that timed out (a rank stalled for > 2 s: the run shares its box) invalidates the region: it is
    # then measured again, at most twice, and the line says so (`handoff_retries`).
    handoff_retries = 0
    while True:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        ev0.record(stream)
        r.start_after(ev0)
        for k in range(args.steps):
            last_frame = step()
        t_enqueue = time.perf_counter() - wall0           # host time to enqueue the K frames (diagnostic: host- vs device-bound)
        r.wait()                                          # the main stream waits for every frame (incl. gather / untile)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - wall0
        try:
            r.check_sync(" (timed region)")
            break
        except RuntimeError as ex:
            handoff_retries += 1
            print("[bench] rank %d: %s" % (rank, ex), file=sys.stderr, flush=True)
            if handoff_retries > 2:
                raise
            dist.barrier()
            r.clear_sync_errors()
            dist.barrier()
    if world > 1:
        # the frame the timed region ended with is still the single-GPU frame
        ok = torch.tensor([1 if (rank != 0 or torch.equal(last_frame, ref_frame)) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        frame_identical["last_timed_frame"] = bool(ok.item())
        if not ok.item():
            raise SystemExit("bench: the last frame of the timed region differs from the single-GPU frame")
    dev_ms = ev0.elapsed_time(ev1)
    print("[bench] rank %d: %.3f ms per step on this rank" % (rank, dev_ms / args.steps), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    clocks = sampler.stop(wall0, wall0 + wall) if rank == 0 else None
    ms_per_step = dev_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the reference-facing call with host buffers (H2D frame state + D2H pixels inside the timed region) ----
    e2e = None
    h2d = 8 * (14 + 6 * len(lights) + 3)                  # rt_camera + lights + object position
    if args.no_e2e:
        e2e = {"value": None, "unit": "Mrays/s", "ms_per_step": None}
    elif world == 1:
        # The reference's master keeps several orders in flight (one goroutine per order, pool.go:161-169; frames are
        # pipelined, master/main.go:265), so e2e is measured the same way: T host threads, each with its own registered
        # scene handle and its own pinned result buffer, each issuing blocking rt_bulk_trace calls (frame state H2D as
        # kernel parameters + K2/K4/K5 + D2H of the 24.9 MB region inside every call).  The latency of one isolated
        # call is reported next to it.
        import ctypes
        from concurrent.futures import ThreadPoolExecutor
        L = rt._lib.load()
        nbytes = W * H * 3
        T = max(1, min(args.e2e_callers, REPLICAS))
        pins = [L.rt_host_alloc(nbytes) for _ in range(T)]
        outs = [np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p)) for p in pins]
        # ONE registered scene (one upload, one BVH); the T concurrent calls each take a lane of it
        os.environ["RT_B200_LANES"] = str(T + 1)
        one = rt.GpuScene(sc_r, device=local_rank, precision=rt.FP32_FAST)
        handles = [one] * T
        e2e_steps = min(args.steps, 100)
        for k in range(T):
            for _ in range(max(1, args.warmup // T)):
                handles[k].bulk_trace(cam13, lights, W, H, 0, 0, W, H, out=outs[k], tan_half_fov=tan_half)
        # (a) latency of a single blocking call, cold L2
        t_single = []
        for _ in range(min(e2e_steps, 20)):
            flush.zero_()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            handles[0].bulk_trace(cam13, lights, W, H, 0, 0, W, H, out=outs[0], tan_half_fov=tan_half)   # blocks until the bytes are in `out`
            t_single.append(time.perf_counter() - t0)
        # (b) throughput with T calls in flight
        per_thread = (e2e_steps + T - 1) // T

        def caller(k):
            for _ in range(per_thread):
                handles[k].bulk_trace(cam13, lights, W, H, 0, 0, W, H, out=outs[k], tan_half_fov=tan_half)
        torch.cuda.synchronize(dev)
        pool = ThreadPoolExecutor(T)
        list(pool.map(caller, range(T)))                  # untimed: every lane allocates its scratch on first use
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        list(pool.map(caller, range(T)))
        e2e_wall = time.perf_counter() - t0
        pool.shutdown()
        n_calls = per_thread * T
        checksum = int(outs[0][::4097].astype(np.uint64).sum())
        e2e_ms = 1e3 * e2e_wall / n_calls
        e2e = {"value": rays_frame / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": nbytes,
               "ms_per_step": e2e_ms, "calls_in_flight": T, "calls": n_calls,
               "single_call_ms": 1e3 * sum(t_single) / len(t_single),
               "single_call_value": rays_frame / (sum(t_single) / len(t_single)) / 1e6,
               "call": "rt_bulk_trace (what Tracer.BulkTrace issues per WorkOrder) on ONE registered scene handle, %d caller threads, pinned host result buffers" % T,
               "checksum": checksum}
        # (c) SURVEY s8 f2: the same calls answered as the protobuf wire bytes of comms.TraceResults (encoded on the GPU,
        # one D2H copy of exactly the message); beside it the C restatement of the encoder on one host core.
        try:
            cap = int(L.rt_trace_results_max_bytes(W * H))
            wpins = [L.rt_host_alloc(cap) for _ in range(T)]
            wouts = [np.ctypeslib.as_array((ctypes.c_uint8 * cap).from_address(p)) for p in wpins]
            wire_len = [0] * T

            def wcaller(k):
                for _ in range(max(1, per_thread // 2)):
                    wire_len[k] = handles[k].bulk_trace_encoded(cam13, lights, W, H, 0, 0, W, H, out=wouts[k], tan_half_fov=tan_half).size
            pool = ThreadPoolExecutor(T)
            list(pool.map(wcaller, range(T)))
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            list(pool.map(wcaller, range(T)))
            wire_wall = time.perf_counter() - t0
            pool.shutdown()
            from oracle import oracle as orc
            t0 = time.perf_counter()
            ref_wire = orc.encode_trace_results(outs[0].reshape(-1, 3))
            cpu_enc_s = time.perf_counter() - t0
            wire_ms = 1e3 * wire_wall / (max(1, per_thread // 2) * T)
            e2e["wire"] = {"call": "rt_bulk_trace_encoded: render + GPU protobuf encode of comms.TraceResults + D2H of the message, %d callers" % T,
                           "ms_per_step": wire_ms, "value": rays_frame / (wire_ms * 1e-3) / 1e6, "message_bytes": int(wire_len[0]),
                           "identical_to_cpu_encoder": bool(ref_wire.size == wire_len[0] and np.array_equal(ref_wire, wouts[0][:wire_len[0]])),
                           "cpu_port_encode_ms_1_core": 1e3 * cpu_enc_s}
            for p in wpins:
                L.rt_host_free(p)
        except Exception as ex:                            # the wire leg is additional evidence, never a reason to lose the line
            e2e["wire"] = {"error": repr(ex)}
        one.close()
        for p in pins:
            L.rt_host_free(p)
    else:
        # N>1: every step = frame state to every rank (kernel parameters) + tiles rendered on every rank + the frame in rank 0's
        # HOST memory.  --e2e-mode host (default): the framebuffers live in pinned host shared memory and every GPU moves its
        # stripes there over its own PCIe link (N links carry the 24.9 MB frame); --e2e-mode peer: NVLink push into rank 0's
        # device framebuffer + one D2H copy over rank 0's link (52 GB/s: 0.48 ms per frame at every N).  Frames are pipelined
        # like `value`; the latency of one isolated frame is reported next to it.
        nbytes = W * H * 3
        e2e_steps = min(args.steps, 200)
        re = r
        if args.e2e_mode == "host":
            try:
                re = parallel.TileShardedRenderer(sc_r, W, H, precision=rt.FP32_FAST, tile=E2E_TILE, device_index=local_rank,
                                                  replicas=REPLICAS, layout_flags=rt.LAYOUT_COLMAJOR, mode="host", shared_scene=args.shared_scene)
            except RuntimeError as ex:                    # raised on every rank alike (e.g. /dev/shm too small): one link it is
                print("[bench] rank %d: host-mode e2e unavailable (%s); using --e2e-mode peer" % (rank, ex), file=sys.stderr, flush=True)
                args.e2e_mode = "peer"
        for _ in range(2 * REPLICAS):                       # lane priming + first touch of the host buffers (untimed)
            hf = re.render_frame(frame_state, flags=render_flags, to_host=True)
        re.wait()
        torch.cuda.synchronize(dev)
        dist.barrier()
        ok = torch.tensor([1 if (rank != 0 or torch.equal(hf, ref_frame.cpu())) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        frame_identical["e2e_path_" + args.e2e_mode] = bool(ok.item())
        if not ok.item():
            raise SystemExit("bench: the %d-GPU host frame differs from the single-GPU frame (e2e mode %s)" % (world, args.e2e_mode))
        t_single = []
        for it in range(3 + min(e2e_steps, 10)):
            flush.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            re.render_frame(frame_state, flags=render_flags, to_host=True)
            re.wait()
            torch.cuda.synchronize(dev)
            dist.barrier()
            if it >= 3:
                t_single.append(time.perf_counter() - t0)
        for attempt in range(3):                            # a timed-out hand-off (a rank stalled > 2 s) => measure again
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            for it in range(e2e_steps):
                hf = re.render_frame(frame_state, flags=render_flags, to_host=True)
            re.wait()
            torch.cuda.synchronize(dev)
            dist.barrier()
            tt = torch.tensor([time.perf_counter() - t0, sum(t_single) / len(t_single)], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            try:
                re.check_sync(" (e2e)")
                break
            except RuntimeError as ex:
                print("[bench] rank %d: %s" % (rank, ex), file=sys.stderr, flush=True)
                if attempt == 2:
                    raise
                handoff_retries += 1
                dist.barrier()
                re.clear_sync_errors()
                dist.barrier()
        e2e_ms = 1e3 * float(tt[0].item()) / e2e_steps
        e2e = {"value": rays_frame / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_ms, "single_call_ms": 1e3 * float(tt[1].item()),
               "checksum": int(hf[::4097].to(torch.int64).sum().item()) if rank == 0 else None, "mode": args.e2e_mode,
               "tile": list(E2E_TILE) if args.e2e_mode == "host" else list(TILE),
               "call": ("TileShardedRenderer.render_frame, mode host: pipelined frames; every rank renders its 8-column stripes in place "
                        "and moves them with one strided copy-engine transfer into the pinned host-shared framebuffer over its own "
                        "PCIe link; hand-off flags in the same host memory"
                        if args.e2e_mode == "host" else
                        "TileShardedRenderer.render_frame(to_host=True): pipelined frames, framebuffer D2H into pinned host memory on rank 0")}
        if re is not r:
            dist.barrier()
            re.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes / its CUDA-event duration) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    avg = {k: sum(v) / max(1, len(v)) for k, v in kern_ms.items()}
    nb, tb = cnt["node_bytes"], cnt["tri_bytes"]
    alg = {
        # K2: node + triangle records fetched by primary rays + one 16 B HitRec per hit
        "primary": cnt["nodes_primary"] * nb + cnt["tris_primary"] * tb + hits_local * 16,
        # K4: shadow-ray records + per hit: HitRec 16, face verts 72, face normals 72, material idx 4 + material 80, colour 24 + lit 4
        "shade": cnt["nodes_shade"] * nb + cnt["tris_shade"] * tb + hits_local * (16 + 72 + 72 + 4 + 80 + 24 + 4),
        # K5: per hit HitRec 16 + colour 24 + lit 4 read, 3 bytes written
        "write": hits_local * (16 + 24 + 4 + 3),
    }
    # fused frame kernel: the same records, no hit list / colour buffer in between, 3 bytes written per hit pixel
    alg["frame"] = (cnt["nodes_primary"] + cnt["nodes_shade"]) * nb + (cnt["tris_primary"] + cnt["tris_shade"]) * tb + hits_local * (72 + 72 + 4 + 80 + 3)
    dom = "frame" if fused_default else max(("primary", "shade", "write"), key=lambda k: avg[k])
    kernel_name = {"frame": "k_frame_fused", "primary": "k_primary_fast_" + ("persistent" if args.persistent else "simple"), "shade": "k_shade",
                   "write": "k_write"}[dom]
    achieved = alg[dom] / (avg[dom] * 1e-3) / 1e9 if avg[dom] > 0 else 0.0
    # what ncu measured for this kernel on this workload (profiles/traffic.json, filled from the committed ncu captures): DRAM and L2
    # bytes per launch, issue-slot and L1-pipe utilisation
    physical = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            physical = json.load(open(tpath)).get(CONFIG, {}).get(kernel_name)
        except Exception:
            physical = None
    traffic = physical.get("dram_bytes") if physical else None
    dram_frac = (traffic / (avg[dom] * 1e-3) / 1e9 / peak) if traffic else None
    # Which roof binds?  If the kernel's measured DRAM traffic moved at more than half of the HBM peak it is an HBM kernel and `frac`
    # is measured DRAM bytes / time / peak; otherwise the working set is cache-resident and the byte figure is algorithmic (records
    # the rays fetch, served by L1/L2): the kernel is then bound by the latency of dependent loads, and `frac` only says how the
    # algorithmic byte rate compares with HBM -- the physical numbers are in `physical`.
    hbm_bound = dram_frac is not None and dram_frac >= 0.5
    roofline = {"bound": "hbm" if hbm_bound else "latency (dependent L1/L2 loads of the traversal; DRAM is idle)",
                "achieved": (traffic / (avg[dom] * 1e-3) / 1e9) if hbm_bound else achieved, "peak": peak, "unit": "GB/s",
                "frac": dram_frac if hbm_bound else achieved / peak, "frac_numerator": "measured DRAM bytes (ncu)" if hbm_bound else "algorithmic bytes",
                "traffic": traffic, "dram_frac_of_peak": dram_frac, "physical": physical,
                "kernel": kernel_name, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": avg[dom],
                "kernels_ms": avg, "algorithmic_bytes": alg,
                "note": "algorithmic bytes = node (112 B) + triangle (64 B) records the rays fetch, from the counting build of the kernels; "
                        "`traffic` / `physical` are ncu's DRAM and L2 bytes, issue-slot and L1-pipe utilisation of one launch of the same kernel "
                        "(profiles/r02_ncu_*_summary.txt).  On C3 the scene is cache-resident (DRAM ~3 % of peak): the limiter is the dependent "
                        "L1/L2 load latency of a BVH visit at the occupancy the registers allow, not bytes and not issue slots",
                "per_ray": {"nodes_primary": cnt["nodes_primary"] / max(1, prim_local), "tris_primary": cnt["tris_primary"] / max(1, prim_local),
                            "nodes_shadow": cnt["nodes_shade"] / max(1, cnt["shadow_rays"]), "tris_shadow": cnt["tris_shade"] / max(1, cnt["shadow_rays"])}}

    # flop side of the traversal roofline: F_alg / measured FMA peak, for the record (the path is bytes-bound)
    import ctypes
    fp32_peak, fp64_peak = ctypes.c_double(0), ctypes.c_double(0)
    L = rt._lib.load()
    L.rt_measure_fma_peak(local_rank, 0, ctypes.byref(fp32_peak))
    L.rt_measure_fma_peak(local_rank, 1, ctypes.byref(fp64_peak))
    F_BOX, F_TRI, F_SHADE_HIT, F_SHADE_LIGHT = 48, 60, 250, 150       # flops per node visit / triangle test / hit / (hit, light)
    flops32 = (cnt["nodes_primary"] + cnt["nodes_shade"]) * F_BOX + (cnt["tris_primary"] + cnt["tris_shade"]) * F_TRI
    flops64 = hits_local * (F_SHADE_HIT + F_SHADE_LIGHT * len(lights))
    roofline["fma_peak_tflops_measured"] = {"fp32": fp32_peak.value, "fp64": fp64_peak.value}
    roofline["flop_side"] = {"fp32_flops": flops32, "fp64_flops": flops64,
                             "t_flops_ms": 1e3 * (flops32 / max(fp32_peak.value, 1e-9) + flops64 / max(fp64_peak.value, 1e-9)) / 1e12,
                             "t_bytes_ms": 1e3 * alg["frame"] / (peak * 1e9),
                             "t_measured_ms": avg["frame"] if fused_default else avg["primary"] + avg["shade"] + avg["write"]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and CONFIG == "c4":
        cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": 1, "kind": "port",
                        "sample": "not timed for C4: the reference-shaped R-tree of 10 M faces alone takes minutes to insert on one core; "
                                  "see the C3 line (python bench.py) for the CPU arm"}
    elif world == 1 and not args.no_cpu_baseline:
        v, rays, dt, stride = oracle_cpu_rate(sc, n_threads=1, budget_s=15.0)
        cpu_baseline = {"value": v, "unit": "Mrays/s", "cores": 1, "kind": "port",
                        "sample": "every %d-th band of 16 pixel columns of the same 4K frame (%d rays, %.1f s), single thread like worker/sequential, "
                                  "reference-shaped R-tree walk (C restatement, not the Go binary)" % (stride, rays, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(
            workload=WORKLOAD, rays_per_frame=rays_frame, hit_pixels=hits_frame, frames_per_s=1e3 / ms_per_step,
            l2=("inputs larger than L2: the resident scene (%s) against 126 MB of L2 and %d frames are in flight on it, one stream each; "
                "no flush inside the pipelined region (the top of the BVH stays cached across frames, as it does in a running worker); the "
                "per-kernel times of `roofline` come from a serial pass with a 256 MiB memset before each frame"
                % ("~300 MB: nodes 32 + 37, triangles 63, fp64 vertices/normals 143" if CONFIG == "c3" else "~2.8 GB", REPLICAS))
            if args.shared_scene else
            ("frames cycle over %d resident replicas of the scene per GPU, one stream each; no flush inside the pipelined region; the "
             "per-kernel times of `roofline` come from a serial pass with a 256 MiB memset before each frame" % REPLICAS),
            scene_copies=1 if args.shared_scene else REPLICAS, scene_create_ms=1e3 * t_create, frames_in_flight=REPLICAS,
            host_enqueue_ms_per_step=1e3 * t_enqueue / args.steps, precision="FP32_FAST traversal, fp64 shading",
            parallelism=("screen tiles 32x16 round-robin over %d rank(s)" % world) + ("" if world == 1 else
                        {"peer": "; every rank pushes its tiles into rank 0's framebuffer over NVLink peer memory (CUDA IPC, 16-byte stores) and "
                                 "the push kernel hands the frame over through arrive/consumed flags in rank 0's memory (no collective)",
                         "peer_nccl": "; tiles pushed into rank 0's framebuffer over NVLink peer memory, 4-byte all_reduce as barrier",
                         "gather": "; NCCL all_gather of packed tiles + de-interleave on rank 0"}[args.mode]),
            bvh=bvh, render_kernel="persistent K2 + K4 + K5" if args.persistent else ("K2 + K4 + K5" if args.pipeline else "k_frame_fused"),
            note="frame_identical_to_single_gpu is checked before the timed region at N > 1"),
        "frame_identical_to_single_gpu": (all(frame_identical.values()) if frame_identical else None), "frame_identity_checks": frame_identical, "handoff_retries": handoff_retries,
        "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "clocks": clocks, "wall_s_timed_region": wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
