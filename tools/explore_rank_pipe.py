"""Single-GPU emulation of ONE rank of an N-rank tile-sharded job, pipelined over R frames in flight: ms per frame for
replicated scenes vs one shared scene (lanes), tile sizes, and R.  (What limits the per-rank frame rate at N = 8?)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import parallel

sc = bench.build_workload()
W, H = bench.W, bench.H
N = int(os.environ.get("EMU_N", 8))
fs = rt.FrameState(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
dev = torch.device("cuda", 0)
os.environ["RT_B200_LANES"] = "17"
cfgs = [(8, False, 32, 16), (8, True, 32, 16), (16, True, 32, 16), (8, True, 64, 32), (8, True, 128, 32), (8, True, 16, 8), (4, True, 32, 16)]
for (R, shared, tw, th) in cfgs:
    be = parallel.GpuTileBackend(sc, 0, rt.FP32_FAST, R, shared)
    per = rt.tiles_per_rank(W, H, tw, th, N)
    bufs = [torch.zeros(per * tw * th * 3, dtype=torch.uint8, device=dev) for _ in range(R)]
    res = {}
    for rank in (0, 3):
        def frame(k):
            be.set_frame_state(k, fs)
            be.render_tiles(k, W, H, tw, th, rank, N, bufs[k])
        for it in range(3 * R):
            frame(it % R)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for s in be.streams:
            s.wait_event(e0)
        K = 400
        for it in range(K):
            frame(it % R)
        for s in be.streams:
            ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        res[rank] = e0.elapsed_time(e1) / K
    print(json.dumps({"N": N, "R": R, "shared": shared, "tile": [tw, th], "ms_per_frame_rank0": res[0], "ms_per_frame_rank3": res[3],
                      "ideal": 0.801 / N}), flush=True)
    for g in set(be.scenes):
        g.close()
    del be, bufs
    torch.cuda.empty_cache()
