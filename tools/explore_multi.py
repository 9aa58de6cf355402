"""Multi-GPU configuration sweep inside ONE torchrun launch (scene built and broadcast once): ms per frame of the pipelined
tile-sharded renderer for (mode, shared scene, tile size, frames in flight, steps).
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 tools/explore_multi.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    import bench
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import parallel
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["RT_B200_LANES"] = "17"
    dist.init_process_group("nccl", device_id=dev)
    sc = bench.build_workload() if rank == 0 else None
    fr = [(sc["camera"], sc["lights"], sc["tan_half_fov"])] if rank == 0 else [None]
    dist.broadcast_object_list(fr, src=0)
    cam, lights, th = fr[0]
    scr = parallel.broadcast_scene(sc, src=0, device=dev)
    W, H = bench.W, bench.H
    fs = rt.FrameState(cam, lights, tan_half_fov=th)
    main_st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main_st)
    cfgs = json.loads(os.environ.get("EXPLORE_CFGS", "[]")) or [
        ["peer", False, 32, 16, 8, 200], ["peer", True, 32, 16, 8, 200], ["peer", True, 64, 32, 8, 200], ["peer", True, 64, 32, 8, 1000],
        ["peer_nccl", True, 64, 32, 8, 200], ["peer", True, 64, 32, 12, 1000], ["peer", True, 128, 32, 8, 1000]]
    for (mode, shared, tw, tth, R, K) in cfgs:
        r = parallel.TileShardedRenderer(scr, W, H, precision=rt.FP32_FAST, tile=(tw, tth), device_index=lr, replicas=R,
                                         layout_flags=rt.LAYOUT_COLMAJOR, mode=mode, shared_scene=shared)
        for _ in range(3 * R):
            r.render_frame(fs)
        r.wait()
        torch.cuda.synchronize(dev)
        res = []
        for rep in range(2):
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(main_st)
            r.start_after(e0)
            for _ in range(K):
                r.render_frame(fs)
            t_enq = time.perf_counter() - t0
            r.wait()
            e1.record(main_st)
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / K, 1e3 * t_enq / K], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res.append([round(float(x), 5) for x in t])
        r.check_sync()
        if rank == 0:
            print(json.dumps({"n": world, "mode": mode, "shared": shared, "tile": [tw, tth], "R": R, "steps": K,
                              "ms_per_frame": [x[0] for x in res], "host_enqueue_ms": [x[1] for x in res]}), flush=True)
        dist.barrier()
        r.close()
        for g in set(r.backend.scenes):
            g.close()
        del r
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
