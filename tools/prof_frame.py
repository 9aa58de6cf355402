"""Profiling driver (GPU only; not part of the product): renders config 3 serially, alternating the three-kernel pipeline and the
fused frame kernel, so that `ncu -k regex:"k_shade|k_primary_fast_simple|k_frame_fused" -s 6 -c 3` captures one launch of each
after two warm-up rounds.   python tools/prof_frame.py [rounds] [c3|c4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import ab_r02  # noqa: E402
import distributed_raytracer_b200 as rt  # noqa: E402

if __name__ == "__main__":
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    which = sys.argv[2] if len(sys.argv) > 2 else "c3"
    W, H = 3840, 2160
    if which == "c4":
        import run_configs
        sc = run_configs.soup_scene(int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000)
    else:
        sc = ab_r02.load_workload()
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    g.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    fb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for r in range(rounds):
        for flags in (rt.FLAG_PIPELINE, rt.FLAG_FUSED):
            flush.zero_()
            torch.cuda.synchronize()
            g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, flags)
            st = g.stats()
            print("round %d %s: primary %.3f shade %.3f write %.3f total %.3f ms" % (r, "pipeline" if flags == rt.FLAG_PIPELINE else "fused",
                                                                                     st["ms_primary"], st["ms_shade"], st["ms_write"], st["ms_total"]), flush=True)
    print("checksum", int(fb[::997].to(torch.int64).sum().item()), "bvh", g.bvh_info())
    g.close()
