"""Why does a rank's 1/8 share of the tiles take much more than 1/8 of the frame?  Region mode vs tile mode at equal work."""
import os, sys
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
fb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")

def run(name, fn):
    ms = []
    for rep in range(6):
        flush.zero_(); torch.cuda.synchronize()
        fn()
        st = g.stats()
        ms.append((st["ms_primary"], st["ms_shade"], st["ms_write"], st["ms_total"], st["hit_pixels"], st["primary_rays"]))
    m = np.median(np.array(ms[2:]), axis=0)
    print("%-44s primary %.3f shade %.3f total %.3f hits %7d rays %8d | us per 1k hits: prim %.2f shade %.2f" % (
        name, m[0], m[1], m[3], m[4], m[5], 1e3 * m[0] / max(m[4] / 1e3, 1e-9), 1e3 * m[1] / max(m[4] / 1e3, 1e-9)), flush=True)

run("full", lambda: g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, 0))
for k in range(0, 8, 2):
    run("band 3840x130 y=%d" % (560 + 130 * k), lambda: g.render_region_device(W, H, 0, 560 + 130 * k, 3840, 130, fb.data_ptr(), None, 0))
run("bbox strip 155x1040 x=1850", lambda: g.render_region_device(W, H, 1850, 560, 155, 1040, fb.data_ptr(), None, 0))
run("bbox strip 310x1040 x=1800", lambda: g.render_region_device(W, H, 1800, 560, 310, 1040, fb.data_ptr(), None, 0))
for N in (8, 16, 32):
    for (tw, th) in ((32, 16), (256, 128)):
        per = rt.tiles_per_rank(W, H, tw, th, N)
        buf = torch.empty(per * tw * th * 3, dtype=torch.uint8, device="cuda")
        run("tiles %dx%d rank0/%d" % (tw, th, N), lambda: g.render_tiles_device(W, H, tw, th, 0, N, buf.data_ptr(), None, 0))
