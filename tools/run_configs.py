"""BASELINE.json configs 4 and 5, one JSON line each.

  C4: python [-m torch.distributed.run ...] tools/run_configs.py c4 [--tris 10000000]
      10 M-triangle random soup (no vertex normals => flat-normal path, default material), 8 white lights on a ring, 4K,
      FP32_FAST, tile-sharded over the ranks.  Reports BVH build time, Mrays/s, and an oracle check on a pixel sample.
  C5: python tools/run_configs.py c5 [--gpus N] [--frames 120]
      120-frame orbit (3 deg per frame) of the subdivided Suzanne, dispatched the way master/pool does: the screen is cut by
      partition() (master/main.go:54-91) into one rectangle per worker, every worker (= one GPU, one host thread standing in
      for the goroutine of pool.Assign, pool.go:148-197) answers Tracer.BulkTrace for its rectangle into HOST memory, the
      master thread assembles the frame.  No gRPC/protobuf (the Go master is not runnable here): this measures the
      worker path + result hand-off, not the reference's wire format.
"""
import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

W, H = 3840, 2160


def soup_scene(n_tris):
    from distributed_raytracer_b200 import host, workloads
    env = host.Environment(os.path.join(ROOT, "tests", "golden", "example", "scene.json"))
    sc = env.scene_arrays()          # example camera + object position (1,1,-1)
    env.close()
    sc["meshes"] = [workloads.triangle_soup(n_tris)]
    sc["lights"] = workloads.ring_lights(8)
    return sc


def oracle_sample(sc, rgb_colmajor, stride, phase, threads=0):
    """Compares a strided pixel sample of a column-major GPU frame with the oracle (conservative-BVH mode)."""
    from oracle import oracle as orc
    S = orc.Scene(sc)
    S.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    t0 = time.time()
    rgb_o, _, face_o, st = S.render_region(W, H, mode=orc.CULLED, step=stride, phase=phase, n_threads=threads, want_colour=False)
    idx = np.arange(W * H) % stride == phase
    a, b = rgb_colmajor.reshape(-1, 3)[idx].astype(np.int32), rgb_o[idx].astype(np.int32)
    within = float((np.abs(a - b).max(axis=1) <= 1).mean())
    exact = float((a == b).all(axis=1).mean())
    S.close()
    return {"sample_pixels": int(idx.sum()), "sample_hits": int(st["hit_pixels"]), "within_1_lsb": within, "identical": exact,
            "oracle_seconds": time.time() - t0}


def run_c4(args):
    import torch
    import torch.distributed as dist
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import parallel
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    sc = soup_scene(args.tris) if rank == 0 else None
    t_gen = time.time() - t0
    if world > 1:
        fr = [(sc["camera"], sc["lights"], sc["tan_half_fov"])] if rank == 0 else [None]
        dist.broadcast_object_list(fr, src=0)
        cam, lights, th = fr[0]
        t0 = time.time()
        scr = parallel.broadcast_scene(sc, src=0, device=dev)
        torch.cuda.synchronize()
        t_bcast = time.time() - t0
    else:
        cam, lights, th, scr, t_bcast = sc["camera"], sc["lights"], sc["tan_half_fov"], sc, 0.0
    main = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main)
    t0 = time.time()
    r = parallel.TileShardedRenderer(scr, W, H, precision=rt.FP32_FAST, tile=(32, 16), device_index=lr, replicas=args.replicas,
                                     layout_flags=rt.LAYOUT_COLMAJOR)
    t_create = time.time() - t0
    g = r.backend.scenes[0]
    bvh = g.bvh_info()
    frame = rt.FrameState(cam, lights, tan_half_fov=th)
    for _ in range(4):
        fb = r.render_frame(frame)
    r.wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    r.start_after(e0)
    for _ in range(args.steps):
        fb = r.render_frame(frame)
    r.wait()
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    # hits: counting pass of this rank's tiles
    g.set_frame(cam, lights, tan_half_fov=th)
    scratch = torch.empty(rt.tiles_per_rank(W, H, 32, 16, world) * 32 * 16 * 3, dtype=torch.uint8, device=dev)
    g.render_tiles_device(W, H, 32, 16, rank, world, scratch.data_ptr(), None, rt.FLAG_COUNT)
    st = g.stats()
    hits = torch.tensor([float(st["hit_pixels"]), float(st["nodes_fetched"]), float(st["tris_fetched"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(hits)
    if rank == 0:
        hits_n = int(hits[0].item())
        rays = W * H + hits_n * len(lights)
        out = {"config": "C4: %d-triangle random soup, 8 lights, 3840x2160, FP32_FAST" % args.tris, "n_gpus": world,
               "ms_per_frame": ms, "mrays_per_s": rays / ms / 1e3, "frames_per_s": 1e3 / ms, "rays_per_frame": rays, "hit_pixels": hits_n,
               "bvh": bvh, "scene_create_s_per_replica": t_create / args.replicas, "scene_generate_s": t_gen, "scene_broadcast_s": t_bcast,
               "nodes_per_ray": hits[1].item() / rays, "tris_per_ray": hits[2].item() / rays, "frames_in_flight": args.replicas}
        if args.check:
            torch.cuda.synchronize()
            out["oracle_check"] = oracle_sample(sc, fb.cpu().numpy(), args.check, 3)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args):
    import ctypes
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import host, workloads
    import bench
    n = args.gpus
    sc = bench.build_workload()
    L = rt._lib.load()
    rects = host.partition(W, H, n)                                   # master/main.go:101 -> :54-91
    assert len(rects) == n, (len(rects), n)
    scenes = [rt.GpuScene(sc, device=d, precision=rt.FP32_FAST) for d in range(n)]       # register(): once per worker
    bufs = []
    for _ in range(n):                                                 # --balanced moves the rectangles every frame: full-size buffers
        p = L.rt_host_alloc(W * H * 3)
        bufs.append(np.ctypeslib.as_array((ctypes.c_uint8 * (W * H * 3)).from_address(p)))
    outs = [bufs[k][:w * h * 3] for k, (x, y, w, h) in enumerate(rects)]
    DS = 8                                                             # the master's cost table is built from every 8th pixel of the last frame

    small = np.zeros((W // DS, H // DS, 3), dtype=np.uint8)           # every 8th pixel of the frame just drawn, filled by the workers

    def next_rects():
        """partition() with the cuts placed by the cost of the frame just drawn (integration/go/master/partition_balanced.go)"""
        r8 = host.partition_balanced(W // DS, H // DS, n, small, args.hit_weight)
        return [(x * DS, y * DS, w * DS, h * DS) for (x, y, w, h) in r8]
    screen = np.zeros((W, H, 3), dtype=np.uint8)                       # column-major like the wire layout (i*H + j)
    centre = np.array([1.0, 1.0, -1.0])
    fov = float(sc["camera"][12])
    tan_half = sc["tan_half_fov"]
    cams = [workloads.orbit_camera(centre, 6.0, np.deg2rad(3.0 * f), 0.0, fov) for f in range(args.frames)]   # 3 deg / frame
    hits = [0] * n

    cur = {"rects": rects}

    def worker(k, cam):
        x, y, w, h = cur["rects"][k]
        out = bufs[k][:w * h * 3]
        scenes[k].bulk_trace(cam, sc["lights"], W, H, x, y, w, h, out=out, tan_half_fov=tan_half)   # Tracer.BulkTrace
        o3 = out.reshape(w, h, 3)
        screen[x:x + w, y:y + h] = o3                                  # master/main.go:166-176 (frame assembly)
        if args.balanced:
            small[x // DS:(x + w) // DS, y // DS:(y + h) // DS] = o3[::DS, ::DS]
        return k

    def frame(cam):
        list(pool.map(lambda k: worker(k, cam), range(n)))             # one frame: all orders in flight, wait for all (coordinator)
        if args.balanced:
            cur["rects"] = next_rects()

    pool = ThreadPoolExecutor(n)
    for cam in cams[:3]:
        frame(cam)
    t_frames = []
    t0 = time.perf_counter()
    for cam in cams:
        tf = time.perf_counter()
        frame(cam)
        t_frames.append(time.perf_counter() - tf)
    wall = time.perf_counter() - t0
    total_hits = 0                                                     # ray accounting: an untimed second pass over the frames
    worst = 0.0
    for cam in cams:
        frame(cam)
        per = [s.stats()["hit_pixels"] for s in scenes]
        total_hits += sum(per)
        worst += max(per) / max(1.0, sum(per) / n)
    rects = cur["rects"]
    rays = W * H * len(cams) + total_hits * len(sc["lights"])
    out = {"config": "C5: %d-frame orbit (3 deg/frame, r=6) of Suzanne x5 subdivided, 3840x2160, 2 lights, FP32_FAST, "
                     "partition() rectangles, one BulkTrace stream per GPU, HOST result buffers + frame assembly" % len(cams),
           "n_gpus": n, "frames": len(cams), "wall_s": wall, "frames_per_s": len(cams) / wall, "mrays_per_s": rays / wall / 1e6,
           "ms_per_frame_median": 1e3 * float(np.median(t_frames)), "ms_per_frame_max": 1e3 * float(np.max(t_frames)),
           "rectangles": rects, "mean_hit_pixels": total_hits / len(cams), "balanced": bool(args.balanced), "hit_weight": args.hit_weight,
           "mean_busiest_worker_hits_over_mean": worst / len(cams),
           "note": "Go master/pool + gRPC not executed (no Go toolchain); threads stand in for pool.Assign goroutines"}
    if args.check:
        sc_last = dict(sc, camera=cams[-1])
        out["oracle_check_last_frame"] = oracle_sample(sc_last, screen.reshape(-1, 3), args.check, 5)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c4", "c5"])
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--replicas", type=int, default=4)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--balanced", action="store_true", help="c5: cuts placed by the cost of the previous frame (SURVEY s8 f4) instead of partition()'s halves")
    ap.add_argument("--hit-weight", type=float, default=30.0, help="c5 --balanced: cost of a hit pixel in units of a miss pixel")
    ap.add_argument("--check", type=int, default=0, help="oracle check on every check-th pixel (0 = off)")
    a = ap.parse_args()
    import __graft_entry__ as ge
    if int(os.environ.get("RANK", 0)) == 0:
        ge.build()
    (run_c4 if a.config == "c4" else run_c5)(a)
