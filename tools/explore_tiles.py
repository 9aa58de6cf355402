"""Single-GPU emulation of the per-rank work of the tile-sharded path: rank r of N renders its tiles; prints kernel times
per (tile size, N, rank) so that load balance, fixed costs and cache effects can be read without N GPUs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import distributed_raytracer_b200 as rt

sc = bench.build_workload()
W, H = bench.W, bench.H
g = rt.GpuScene(sc, precision=rt.FP32_FAST)
g.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (tw, th) in [(32, 16), (64, 32), (128, 64), (256, 128), (8, 4)]:
    for N in (1, 2, 4, 8):
        per = rt.tiles_per_rank(W, H, tw, th, N)
        buf = torch.empty(per * tw * th * 3, dtype=torch.uint8, device="cuda")
        rows = []
        for r in range(N):
            ms = []
            for rep in range(4):
                flush.zero_(); torch.cuda.synchronize()
                g.render_tiles_device(W, H, tw, th, r, N, buf.data_ptr(), None, 0)
                st = g.stats()
                ms.append((st["ms_primary"], st["ms_shade"], st["ms_write"], st["ms_total"], st["hit_pixels"]))
            m = np.median(np.array(ms[1:]), axis=0)
            rows.append(m)
        rows = np.array(rows)
        print(json.dumps({"tile": [tw, th], "N": N, "primary_max": float(rows[:, 0].max()), "shade_max": float(rows[:, 1].max()),
                          "total_max": float(rows[:, 3].max()), "total_mean": float(rows[:, 3].mean()),
                          "hits_max": int(rows[:, 4].max()), "hits_min": int(rows[:, 4].min())}), flush=True)
