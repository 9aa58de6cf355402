"""Multi-GPU correctness check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py

Every mode of TileShardedRenderer (peer = NVLink tile push + device-side hand-off flags, host = every GPU stores its tiles into a
pinned host-shared framebuffer, peer_nccl = NVLink push + NCCL barrier,
gather = NCCL all_gather + de-interleave) must deliver, on rank 0, exactly the frame a single GPU renders with
rt_render_region -- for a sequence of DIFFERENT cameras pipelined over the frame slots (so a stale or mixed slot shows up),
device-resident and through the device->host path.  Prints one 'MULTIGPU OK ...' line on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import distributed_raytracer_b200 as rt
    from distributed_raytracer_b200 import host, parallel, workloads
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    env = host.Environment(os.path.join(ROOT, "tests", "golden", "example", "scene.json"))
    sc = env.scene_arrays()
    env.close()
    sc["meshes"] = [workloads.subdivide(sc["meshes"][0], 2)]
    W, H, S, F = 1280, 720, 3, 11
    cams = [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(17.0 * f), 0.03 * f, 1.04719755) for f in range(F)]
    ref = None
    if rank == 0:
        g = rt.GpuScene(sc, device=lr, precision=rt.FP32_FAST)
        ref = []
        for c in cams:
            g.set_frame(c, sc["lights"])
            ref.append(g.render_region(W, H))
        g.close()
    checked = 0
    for mode in ("peer", "host", "peer_nccl", "gather"):
        for to_host in (False, True):
            if to_host and mode in ("gather", "host"):
                continue
            r = parallel.TileShardedRenderer(sc, W, H, precision=rt.FP32_FAST, tile=(32, 16), device_index=lr, replicas=S,
                                             layout_flags=rt.LAYOUT_COLMAJOR, mode=mode)
            got = []
            for f0 in range(0, F, S):                       # S frames in flight, then collect (a slot holds one frame)
                outs = []
                for f in range(f0, min(F, f0 + S)):
                    fs = rt.FrameState(cams[f], sc["lights"])
                    if mode == "gather":
                        outs.append(r.render_frame(cams[f], sc["lights"]))
                    else:
                        outs.append(r.render_frame(fs, to_host=to_host))
                r.wait()
                torch.cuda.synchronize(dev)
                if rank == 0:
                    got += [o.cpu().numpy().reshape(-1, 3).copy() for o in outs]
                if not to_host:
                    dist.barrier()      # device-resident frames: a slot may be overwritten (by ANY rank) once the frame S later is enqueued
            # back-to-back frames without collecting in between: only the last S are still resident
            for f in range(F):
                last = r.render_frame(rt.FrameState(cams[f], sc["lights"]), to_host=to_host) if mode != "gather" else \
                    r.render_frame(cams[f], sc["lights"])
            r.wait()
            torch.cuda.synchronize(dev)
            r.check_sync()
            if rank == 0:
                for f in range(F):
                    assert np.array_equal(got[f], ref[f]), (mode, to_host, f, int((got[f] != ref[f]).any(axis=1).sum()))
                assert np.array_equal(last.cpu().numpy().reshape(-1, 3), ref[F - 1]), (mode, to_host, "back-to-back")
                checked += F + 1
            dist.barrier()
            r.close()
            del r
    if rank == 0:
        print("MULTIGPU OK world=%d frames_checked=%d" % (world, checked), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
