"""Round-2 kernel A/B on config 3 (GPU only; not part of the product).  One subprocess per variant so that the environment
knobs (read at rt_scene_create) and RT_B200_LIB (another build of the library) can differ; one JSON line per variant:

  serial  : per-kernel CUDA-event times of single frames with a 256 MiB memset before each (cold L2)
  piped   : ms per frame with 8 frames in flight on one resident scene (what bench.py's `value` measures)

  python tools/ab_r02.py            # the whole matrix
  python tools/ab_r02.py one NAME   # (internal) run one variant in this process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/c3_workload.npz"

VARIANTS = {
    # name: (env, render flags by name).  build_variants/*.so: see tools/build_variants.sh
    "r01_like": ({"RT_B200_LIB": "build_variants/librt_b200_nooct.so", "RT_B200_CULL": "0", "RT_B200_MAX_LEAF_SHADE": "4", "RT_B200_SHADE_MINB": "8"}, ["PIPELINE"]),
    "pipeline": ({}, ["PIPELINE"]),
    "fused_ld128": ({"RT_B200_LIB": "build_variants/librt_b200_ld128.so"}, []),
    "fused_smemstack": ({"RT_B200_LIB": "build_variants/librt_b200_smemstack.so"}, []),
    "fused_tri48": ({"RT_B200_LIB": "build_variants/librt_b200_tri48.so"}, []),
    "fused_cta64": ({"RT_B200_LIB": "build_variants/librt_b200_cta64.so"}, []),
    "fused": ({}, []),
    "fused_minb8": ({"RT_B200_SHADE_MINB": "8"}, []),
    "fused_minb10": ({"RT_B200_SHADE_MINB": "10"}, []),
    "fused_any4": ({"RT_B200_MAX_LEAF_SHADE": "4"}, []),
    # leaves of any-hit rays postponed (RT_ANY_DEFER, now the default) against the plain loop, and the knobs that may move with it
    "fused_nodefer": ({"RT_B200_LIB": "build_variants/librt_b200_nodefer.so"}, []),
    "pipeline_nodefer": ({"RT_B200_LIB": "build_variants/librt_b200_nodefer.so"}, ["PIPELINE"]),
    "fused_pf1": ({"RT_B200_LIB": "build_variants/librt_b200_pf1.so"}, []),
    "fused_pf2": ({"RT_B200_LIB": "build_variants/librt_b200_pf2.so"}, []),
    # 64-byte quantised nodes (RT_NODEQ, now the default) against the 112-byte float nodes
    "fused_noq": ({"RT_B200_LIB": "build_variants/librt_b200_noq.so"}, []),
    "pipeline_noq": ({"RT_B200_LIB": "build_variants/librt_b200_noq.so"}, ["PIPELINE"]),
    "fused_minb10": ({"RT_B200_SHADE_MINB": "10"}, []),
    # L1 cache-policy hints: triangle records L1::no_allocate, node lines L1::evict_last
    "fused_pair_minb8": ({"RT_B200_LIB": "build_variants/librt_b200_pair.so", "RT_B200_SHADE_MINB": "8"}, []),
    "fused_pair_minb7": ({"RT_B200_LIB": "build_variants/librt_b200_pair.so", "RT_B200_SHADE_MINB": "7"}, []),
    "fused_nopair": ({"RT_B200_LIB": "build_variants/librt_b200_nopair.so"}, []),          # -DRT_LEAF_PAIR=0: triangles of a leaf loaded one after the other
    "pipeline_nopair": ({"RT_B200_LIB": "build_variants/librt_b200_nopair.so"}, ["PIPELINE"]),
    "fused_trialloc": ({"RT_B200_LIB": "build_variants/librt_b200_trialloc.so"}, []),      # -DRT_TRI_NOALLOC=0: triangle records allocate in L1
    "fused_nodeel": ({"RT_B200_LIB": "build_variants/librt_b200_nodeel.so"}, []),          # -DRT_NODE_EVICT_LAST=1
    "fused_hitstream": ({"RT_B200_LIB": "build_variants/librt_b200_hitstream.so"}, []),    # -DRT_HIT_STREAM=1: per-hit fp64 records with ld.global.cs
    # persistent SM-affine frame kernel (k_frame_fused_sm); chunk = 32 8x4 blocks unless the build says otherwise
    "fused_sm": ({"RT_B200_SM_AFFINE": "1"}, []),
    "fused_sm_minb8": ({"RT_B200_SM_AFFINE": "1", "RT_B200_SHADE_MINB": "8"}, []),
    "fused_sm_c16": ({"RT_B200_SM_AFFINE": "1", "RT_B200_LIB": "build_variants/librt_b200_chunk16.so"}, []),
    "fused_sm_c64": ({"RT_B200_SM_AFFINE": "1", "RT_B200_LIB": "build_variants/librt_b200_chunk64.so"}, []),
    "fused_sm_c128": ({"RT_B200_SM_AFFINE": "1", "RT_B200_LIB": "build_variants/librt_b200_chunk128.so"}, []),
    "fused_any3": ({"RT_B200_MAX_LEAF_SHADE": "3"}, []),
    "fused_leaf2": ({"RT_B200_MAX_LEAF": "2"}, []),
    "fused_leaf3": ({"RT_B200_MAX_LEAF": "3"}, []),
    "fused_leaf6": ({"RT_B200_MAX_LEAF": "6"}, []),
    "fused_leaf8": ({"RT_B200_MAX_LEAF": "8"}, []),
    "fused_defer32": ({"RT_B200_ANY_DEFER": "32"}, []),
    "fused_defer12": ({"RT_B200_ANY_DEFER": "12"}, []),
    "fused_any1": ({"RT_B200_MAX_LEAF_SHADE": "1"}, []),
    "fused_any8": ({"RT_B200_MAX_LEAF_SHADE": "8"}, []),
}


def load_workload():
    import numpy as np
    if os.path.exists(CACHE):
        z = np.load(CACHE)
        return {"meshes": [{k: z["m_" + k] for k in ("verts", "normals", "faces", "materials")}], "obj_pos": z["obj_pos"],
                "obj_mesh": z["obj_mesh"], "camera": z["camera"], "lights": z["lights"], "tan_half_fov": float(z["tan_half_fov"])}
    import bench
    sc = bench.build_workload()
    m = sc["meshes"][0]
    np.savez(CACHE, obj_pos=sc["obj_pos"], obj_mesh=sc["obj_mesh"], camera=sc["camera"], lights=sc["lights"],
             tan_half_fov=sc["tan_half_fov"], **{"m_" + k: m[k] for k in ("verts", "normals", "faces", "materials")})
    return sc


def run_one(name):
    import numpy as np
    import torch
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import parallel
    env, flag_names = VARIANTS[name]
    flags = 0
    for f in flag_names:
        flags |= getattr(rt, "FLAG_" + f)
    W, H = 3840, 2160
    sc = load_workload()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    r = parallel.TileShardedRenderer(sc, W, H, precision=rt.FP32_FAST, tile=(32, 16), device_index=0, replicas=8,
                                     layout_flags=rt.LAYOUT_COLMAJOR, mode="peer", shared_scene=True)
    g = r.backend.scenes[0]
    fb = r.frame[0]
    g.set_frame(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, flags | rt.FLAG_COUNT)
    cnt = g.stats()
    ms = {"primary": [], "shade": [], "write": [], "total": []}
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 0.4:
        flush.zero_()
        torch.cuda.synchronize(dev)
    for _ in range(12):
        flush.zero_()
        torch.cuda.synchronize(dev)
        g.render_region_device(W, H, 0, 0, W, H, fb.data_ptr(), None, flags)
        st = g.stats()
        for k in ms:
            ms[k].append(st["ms_" + k])
    serial = {k: float(np.median(v[2:])) for k, v in ms.items()}
    checksum = int(fb[::997].to(torch.int64).sum().item())
    frame_state = rt.FrameState(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
    piped = []
    for rep in range(3):
        for _ in range(16):
            r.render_frame(frame_state, flags=flags)
        r.wait()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        r.start_after(ev0)
        K = 200
        for _ in range(K):
            r.render_frame(frame_state, flags=flags)
        r.wait()
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        piped.append(ev0.elapsed_time(ev1) / K)
    rays = W * H + cnt["shadow_rays"]
    print(json.dumps({"variant": name, "env": env, "flags": flag_names, "serial_ms": serial, "piped_ms": min(piped), "piped_all": piped,
                      "grays_s_piped": rays / min(piped) / 1e6, "hits": cnt["hit_pixels"], "checksum": checksum,
                      "nodes_primary_per_ray": cnt["nodes_primary"] / (W * H), "tris_primary_per_ray": cnt["tris_primary"] / (W * H),
                      "nodes_shadow_per_ray": cnt["nodes_shade"] / max(1, cnt["shadow_rays"]),
                      "tris_shadow_per_ray": cnt["tris_shade"] / max(1, cnt["shadow_rays"]), "bvh": g.bvh_info()}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        run_one(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(VARIANTS)
    load_workload()
    for name in names:
        env = dict(os.environ)
        env.update(VARIANTS[name][0])
        if "RT_B200_LIB" in env:
            env["RT_B200_LIB"] = os.path.join(ROOT, env["RT_B200_LIB"])
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], env=env, capture_output=True, text=True, timeout=600)
        out = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        print(out[-1] if out else json.dumps({"variant": name, "error": (p.stderr or p.stdout)[-600:]}), flush=True)
