"""One rank's share of an N-GPU frame, rendered ALONE on one GPU (GPU only; not part of the product): what a rank of the
tile-sharded pipeline does per frame -- frame state, fused frame kernel over its tiles, push of the packed tiles into a full
framebuffer (here a local one, no hand-off flags) -- with R frames in flight, as bench.py's `value` path runs it.

Prints, per (N, rank): ms per frame in steady state (200 frames) and over 20 frames from an idle pipeline (the driver's
step count: fill and drain inside the timed region), against the ideal 1/N of the N = 1 frame.

  python tools/emulate_rank.py [N ...]            e.g.  python tools/emulate_rank.py 1 2 4 8
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import ab_r02  # noqa: E402
import distributed_raytracer_b200 as rt  # noqa: E402

W, H = 3840, 2160


def measure(g, sc, N, rank, R, tile, steps, flags=0):
    tw, th = tile
    per = rt.tiles_per_rank(W, H, tw, th, N)
    dev = torch.device("cuda", 0)
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(R)]
    tiles = [torch.zeros(per * tw * th * 3, dtype=torch.uint8, device=dev) for _ in range(R)]
    fbs = [torch.zeros(W * H * 3, dtype=torch.uint8, device=dev) for _ in range(R)]
    frame = rt.FrameState(sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])

    def enqueue(k):
        s = k % R
        g.render_push_tiles(frame, W, H, tw, th, rank, N, tiles[s].data_ptr(), fbs[s].data_ptr(), streams[s].cuda_stream,
                            flags | rt.LAYOUT_COLMAJOR | rt.LANE(s + 1))
    for k in range(2 * R):
        enqueue(k)
    torch.cuda.synchronize(dev)
    out = []
    for n_steps in steps:
        best = 1e9
        for rep in range(5):
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for st in streams:
                st.wait_event(e0)
            for k in range(n_steps):
                enqueue(k)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)
            e1.record(main)
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1) / n_steps)
        out.append(best)
    return out


if __name__ == "__main__":
    Ns = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
    R = int(os.environ.get("EMU_R", "12"))
    tile = tuple(int(v) for v in os.environ.get("EMU_TILE", "32x16").split("x"))
    sc = ab_r02.load_workload()
    os.environ["RT_B200_LANES"] = str(R + 1)
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    import time
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.4:
        flush.zero_()
        torch.cuda.synchronize()
    base = None
    for N in Ns:
        for rank in sorted({0, N // 2}):
            steady, short = measure(g, sc, N, rank, R, tile, (200, 20))
            if N == 1:
                base = steady
            print(json.dumps({"N": N, "rank": rank, "R": R, "tile": list(tile), "ms_per_frame_200": steady, "ms_per_frame_20": short,
                              "ideal": (base / N) if base else None, "eff_200": (base / N / steady) if base else None,
                              "eff_20": (base / N / short) if base else None}), flush=True)
    g.close()
