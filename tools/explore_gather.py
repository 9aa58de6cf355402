"""2-GPU (torchrun) timing of the pieces of the tile path: render tiles, all_gather, untile."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import parallel

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sc = bench.build_workload() if rank == 0 else None
fr = [(sc["camera"], sc["lights"], sc["tan_half_fov"])] if rank == 0 else [None]
dist.broadcast_object_list(fr, src=0)
cam, lights, th = fr[0]
scr = parallel.broadcast_scene(sc, src=0, device=dev)
W, H = bench.W, bench.H
r = parallel.TileShardedRenderer(scr, W, H, replicas=1, device_index=lr, layout_flags=rt.LAYOUT_COLMAJOR)
g = r.backend.scenes[0]
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
res = {"render": [], "gather": [], "untile": []}
for it in range(12):
    dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(4)]
    g.set_frame(cam, lights, tan_half_fov=th)
    e[0].record(st)
    g.render_tiles_device(W, H, 32, 16, rank, world, r.local[0].data_ptr(), st.cuda_stream, 0)
    e[1].record(st)
    dist.all_gather_into_tensor(r.gathered[0], r.local[0])
    e[2].record(st)
    if rank == 0:
        g.untile_device(W, H, 32, 16, world, r.gathered[0].data_ptr(), r.frame[0].data_ptr(), st.cuda_stream, rt.LAYOUT_COLMAJOR)
    e[3].record(st)
    torch.cuda.synchronize()
    if it >= 2:
        res["render"].append(e[0].elapsed_time(e[1])); res["gather"].append(e[1].elapsed_time(e[2])); res["untile"].append(e[2].elapsed_time(e[3]))
print("rank", rank, {k: round(float(np.median(v)), 4) for k, v in res.items()}, flush=True)
dist.barrier(); dist.destroy_process_group()
