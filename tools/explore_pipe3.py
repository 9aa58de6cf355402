"""torchrun, N ranks, NCCL: adds the pieces of the peer pipeline one by one to find what costs throughput."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import parallel
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sc = bench.build_workload()
W, H = bench.W, bench.H
R = 4
be = parallel.GpuTileBackend(sc, lr, rt.FP32_FAST, R)
main = torch.cuda.Stream(); torch.cuda.set_stream(main)
tw, th = 32, 16
fb_local = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
# peer buffers
nbytes = W * H * 3
handles = [None]
if rank == 0:
    pb = [rt.DeviceBuffer(lr, nbytes) for _ in range(R)]; handles = [[b.export() for b in pb]]
dist.broadcast_object_list(handles, src=0)
if rank != 0:
    pb = [rt.DeviceBuffer(lr, nbytes, ipc_handle=h) for h in handles[0]]
flag = torch.zeros(1, dtype=torch.int32, device=dev)
dist.all_reduce(flag); torch.cuda.synchronize()
comm = torch.cuda.Stream()
done = [None] * R

def run(name, peer, events, barrier):
    global done
    done = [None] * R
    def frame(k):
        s = k % R
        st = be.streams[s]
        be.set_frame(s, sc["camera"], sc["lights"], tan_half_fov=sc["tan_half_fov"])
        if events and done[s] is not None: st.wait_event(done[s])
        be.render_tiles_peer(s, W, H, tw, th, rank, world, pb[s].ptr if peer else fb_local.data_ptr(), rt.LAYOUT_COLMAJOR)
        if events:
            ev = torch.cuda.Event(); ev.record(st); comm.wait_event(ev)
            with torch.cuda.stream(comm):
                if barrier: dist.all_reduce(flag)
                d = torch.cuda.Event(); d.record(comm)
            done[s] = d
    for k in range(8): frame(k)
    torch.cuda.synchronize(); dist.barrier()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in be.streams: st.wait_event(e0)
    comm.wait_event(e0)
    t0 = time.perf_counter()
    for k in range(K): frame(k)
    th_ = time.perf_counter() - t0
    for st in be.streams:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)
    e = torch.cuda.Event(); e.record(comm); main.wait_event(e)
    e1.record(main); torch.cuda.synchronize()
    print("rank %d/%d %-34s: %.4f ms per rank-frame (host enqueue %.3f)" % (rank, world, name, e0.elapsed_time(e1) / K, 1e3 * th_ / K), flush=True)
    dist.barrier()

run("local fb, no events", False, False, False)
run("local fb, events", False, True, False)
run("local fb, events+barrier", False, True, True)
run("peer fb, no events", True, False, False)
run("peer fb, events", True, True, False)
run("peer fb, events+barrier", True, True, True)
dist.destroy_process_group()
