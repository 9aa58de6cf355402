"""Single GPU, no NCCL: throughput of a rank's share of the tiles when R frames are in flight on R streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import parallel

sc = bench.build_workload()
W, H = bench.W, bench.H
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
be = parallel.GpuTileBackend(sc, 0, rt.FP32_FAST, R)
main = torch.cuda.Stream(); torch.cuda.set_stream(main)
for N in (1, 2, 4, 8):
    for (tw, th) in ((32, 16), (128, 64)):
        per = rt.tiles_per_rank(W, H, tw, th, N)
        bufs = [torch.empty(per * tw * th * 3, dtype=torch.uint8, device="cuda") for _ in range(R)]
        def frame(k):
            s = k % R
            be.set_frame(s, sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
            be.render_tiles(s, W, H, tw, th, 0, N, bufs[s], 0)
        for k in range(8): frame(k)
        torch.cuda.synchronize()
        K = 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in be.streams: st.wait_event(e0)
        for k in range(K): frame(k)
        for st in be.streams:
            e = torch.cuda.Event(); e.record(st); main.wait_event(e)
        e1.record(main); torch.cuda.synchronize()
        print("R=%d N=%d tile %dx%d : %.4f ms per rank-frame  (x%.2f vs 1/N of N=1)" % (R, N, tw, th, e0.elapsed_time(e1) / K, 0), flush=True)
