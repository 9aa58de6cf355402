"""GPU exploration harness (not part of the product): times the render kernels of config 3 under the
build/launch knobs (RT_B200_MAX_LEAF, RT_B200_TOP_SMEM, RT_B200_REFILL, RT_FLAG_SIMPLE) and prints one JSON line per
variant.  Usage: python tools/explore.py [quick]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import distributed_raytracer_b200 as rt  # noqa: E402


def run_variant(sc, name, env, flags=0, reps=5, precision=rt.FP32_FAST, W=bench.W, H=bench.H):
    for k, v in env.items():
        os.environ[k] = str(v)
    t0 = time.time()
    g = rt.GpuScene(sc, precision=precision)
    t_create = time.time() - t0
    fb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, flags | rt.FLAG_COUNT)
    cnt = g.stats()
    ms = {"primary": [], "shade": [], "write": [], "total": []}
    for _ in range(reps + 2):
        flush.zero_()
        torch.cuda.synchronize()
        g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, flags)
        st = g.stats()
        for k in ms:
            ms[k].append(st["ms_" + k])
    med = {k: float(np.median(v[2:])) for k, v in ms.items()}
    rays = W * H + cnt["shadow_rays"]
    out = {"variant": name, "env": env, "flags": flags, "ms": med, "mrays_s": rays / med["total"] / 1e3, "create_s": round(t_create, 3),
           "bvh": g.bvh_info(), "nodes_primary_per_ray": cnt["nodes_primary"] / (W * H), "tris_primary_per_ray": cnt["tris_primary"] / (W * H),
           "nodes_shadow_per_ray": cnt["nodes_shade"] / max(1, cnt["shadow_rays"]), "tris_shadow_per_ray": cnt["tris_shade"] / max(1, cnt["shadow_rays"]),
           "hits": cnt["hit_pixels"], "checksum": int(fb[::997].sum().item())}
    print(json.dumps(out), flush=True)
    g.close()
    return out


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    sc = bench.build_workload()
    base = {"RT_B200_MAX_LEAF": 4, "RT_B200_TOP_SMEM": 0, "RT_B200_REFILL": 1, "RT_B200_SHADE_MINB": 1}
    which = sys.argv[2:] if len(sys.argv) > 2 else ["kernels", "shade", "leaf"]
    run_variant(sc, "default", dict(base))
    if "kernels" in which:
        run_variant(sc, "simple", dict(base), flags=rt.FLAG_SIMPLE)
        run_variant(sc, "top_smem", dict(base, RT_B200_TOP_SMEM=1))
        for r in (8, 16, 24, 32):
            run_variant(sc, "refill%d" % r, dict(base, RT_B200_REFILL=r))
    if "shade" in which:
        for mb in (5, 6, 8):
            run_variant(sc, "shade_minb%d" % mb, dict(base, RT_B200_SHADE_MINB=mb), flags=rt.FLAG_SIMPLE)
    if "leaf" in which:
        for ml in (1, 2, 8):
            run_variant(sc, "maxleaf%d" % ml, dict(base, RT_B200_MAX_LEAF=ml), flags=rt.FLAG_SIMPLE)
    if not quick and "exact" in which:
        run_variant(sc, "exact_c3", dict(base), precision=rt.FP64_EXACT, reps=2)
