"""rt_group_bulk_trace on every GPU of the box (GPU only; not part of the product): ONE process, config 3, the whole 4K frame per
order into pinned host memory, T concurrent callers (the reference worker serves every BulkTrace on its own goroutine).
Checks the frame against the single-GPU rt_bulk_trace byte for byte, then prints one JSON line per T.

  python tools/bench_group.py [calls_per_thread]
"""
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402

import ab_r02  # noqa: E402
import distributed_raytracer_b200 as rt  # noqa: E402

W, H = 3840, 2160

if __name__ == "__main__":
    per_thread = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    sc = ab_r02.load_workload()
    L = rt._lib.load()
    n_dev = L.rt_device_count()
    nbytes = W * H * 3
    one = rt.GpuScene(sc, device=0, precision=rt.FP32_FAST)
    ref = one.bulk_trace(sc["camera"], sc["lights"], W, H, 0, 0, W, H, tan_half_fov=sc["tan_half_fov"]).copy()
    hits = one.stats()["hit_pixels"]
    one.close()
    rays = W * H + hits * len(sc["lights"])
    t0 = time.perf_counter()
    grp = rt.GpuGroup(sc, devices=None, precision=rt.FP32_FAST)
    t_create = time.perf_counter() - t0
    for T in (1, 2, 4, 8):
        pins = [L.rt_host_alloc(nbytes) for _ in range(T)]
        outs = [np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p)).reshape(-1, 3) for p in pins]

        def caller(k, n=per_thread):
            for _ in range(n):
                grp.bulk_trace(sc["camera"], sc["lights"], W, H, 0, 0, W, H, out=outs[k], tan_half_fov=sc["tan_half_fov"])
        pool = ThreadPoolExecutor(T)
        list(pool.map(lambda k: caller(k, 5), range(T)))            # lane priming
        identical = all(np.array_equal(o, ref) for o in outs)
        t0 = time.perf_counter()
        list(pool.map(caller, range(T)))
        wall = time.perf_counter() - t0
        pool.shutdown()
        ms = 1e3 * wall / (T * per_thread)
        print(json.dumps({"devices": n_dev, "callers": T, "ms_per_order": ms, "mrays_per_s": rays / ms / 1e3, "frames_per_s": 1e3 / ms,
                          "identical_to_single_gpu": bool(identical), "group_create_s": t_create, "d2h_bytes_per_order": nbytes,
                          "call": "rt_group_bulk_trace, whole 4K frame per order, pinned host buffers, one process"}), flush=True)
        for p in pins:
            L.rt_host_free(p)
    grp.close()
