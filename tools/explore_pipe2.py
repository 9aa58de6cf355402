"""torchrun, N ranks: each rank renders ITS share with R streams, NO collective at all, optional NCCL init."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import parallel
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
sc = bench.build_workload()
W, H = bench.W, bench.H
R = 4
be = parallel.GpuTileBackend(sc, lr, rt.FP32_FAST, R)
main = torch.cuda.Stream(); torch.cuda.set_stream(main)
tw, th = 32, 16
per = rt.tiles_per_rank(W, H, tw, th, world)
bufs = [torch.empty(per * tw * th * 3, dtype=torch.uint8, device="cuda") for _ in range(R)]
fb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
for mode in ("tiles_device", "tiles_peer_local"):
    def frame(k):
        s = k % R
        be.set_frame(s, sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
        if mode == "tiles_device":
            be.render_tiles(s, W, H, tw, th, rank, world, bufs[s], 0)
        else:
            be.render_tiles_peer(s, W, H, tw, th, rank, world, fb.data_ptr(), 0)
    for k in range(8): frame(k)
    torch.cuda.synchronize(); dist.barrier()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in be.streams: st.wait_event(e0)
    for k in range(K): frame(k)
    for st in be.streams:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)
    e1.record(main); torch.cuda.synchronize()
    print("rank %d/%d %s: %.4f ms per rank-frame" % (rank, world, mode, e0.elapsed_time(e1) / K), flush=True)
    dist.barrier()
dist.destroy_process_group()
