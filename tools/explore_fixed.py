"""Where do the fixed costs of short frames come from?  (a) warm vs cold L2, (b) the all-miss part of the screen."""
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

def run(name, fn, do_flush):
    ms = []
    for rep in range(6):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        fn()
        st = g.stats()
        ms.append((st["ms_primary"], st["ms_shade"], st["ms_write"], st["ms_total"], st["hit_pixels"]))
    m = np.median(np.array(ms[2:]), axis=0)
    print("%-40s flush=%d primary %.3f shade %.3f write %.3f total %.3f hits %d" % (name, do_flush, m[0], m[1], m[2], m[3], m[4]), flush=True)

fb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
for do_flush in (1, 0):
    run("full frame", lambda: g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, 0), do_flush)
    run("hit bbox 1240x1040 @(1300,560)", lambda: g.render_region_device(W, H, 1300, 560, 1240, 1040, fb.data_ptr(), None, 0), do_flush)
    run("all-miss strip 3840x500 @(0,0)", lambda: g.render_region_device(W, H, 0, 0, 3840, 500, fb.data_ptr(), None, 0), do_flush)
    run("centre 512x512", lambda: g.render_region_device(W, H, 1664, 824, 512, 512, fb.data_ptr(), None, 0), do_flush)
    run("centre 128x128", lambda: g.render_region_device(W, H, 1856, 1016, 128, 128, fb.data_ptr(), None, 0), do_flush)
    for N in (2, 8):
        per = rt.tiles_per_rank(W, H, 32, 16, N)
        buf = torch.empty(per * 32 * 16 * 3, dtype=torch.uint8, device="cuda")
        run("tiles 32x16 rank0/%d" % N, lambda: g.render_tiles_device(W, H, 32, 16, 0, N, buf.data_ptr(), None, 0), do_flush)
