import os, sys, json, time
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as ge
ge.build(); ge.smoke()
os.environ["RT_B200_SHADE32"] = "1"
import distributed_raytracer_b200 as rt
from oracle import oracle as orc, scene_loader as sl
import bench, torch
ex = sl.environment_from_file("tests/golden/example/scene.json")
S = orc.Scene(ex); rgb_o = S.render_region(640, 480)[0]
g = rt.GpuScene(ex, precision=rt.FP32_FAST); rgb = g.render_region(640, 480); g.close()
d = np.abs(rgb.astype(int) - rgb_o.astype(int)).max(axis=1)
out = {"shade32_c1_640x480_within1": float((d <= 1).mean()), "gt1": int((d > 1).sum())}
sc = bench.build_workload()
for flag in ("0", "1"):
    os.environ["RT_B200_SHADE32"] = flag
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    g.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    buf = torch.zeros(3840 * 2160 * 3, dtype=torch.uint8, device="cuda")
    ms = []
    for it in range(8):
        g.render_region_device(3840, 2160, 0, 0, 3840, 2160, buf.data_ptr(), None, 0)
        ms.append(g.stats()["ms_shade"])
    out["k4_ms_shade32_" + flag] = float(np.median(ms[2:]))
    out["checksum_" + flag] = int(buf[::4097].to(torch.int64).sum().item())
    g.close()
print(json.dumps(out))
