"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU tile sharding (scene broadcast, interleaved
tile ownership, gather, de-interleave) with the device replaced by the tests/emul replay of the same
device code.  The real NCCL path is exercised by bench.py --gpus N on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import EXAMPLE_SCENE, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class EmulTileBackend:
    """Test backend: same interface as parallel.GpuTileBackend, rendering through tests/emul on the CPU."""

    def __init__(self, scene, device_index, precision):
        import emul_build as eb
        self.E = eb.EmulScene(scene, precision)
        self.cam = None
        self.lights = None

    def set_frame(self, cam13, lights, obj_pos=None, tan_half_fov=float("nan")):
        self.cam, self.lights = np.asarray(cam13), np.asarray(lights)

    pipelined = False

    def n_slots(self):
        return 1

    def alloc(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def render_tiles(self, W, H, tw, th, rank, n_ranks, out, flags=0):
        out.zero_()
        self.E.render_tiles(self.cam, self.lights, W, H, tw, th, rank, n_ranks, out.numpy(), flags)

    def untile(self, W, H, tw, th, n_ranks, all_tiles, out, flags=0):
        from test_emul import numpy_untile
        img = numpy_untile(all_tiles.numpy(), W, H, tw, th, n_ranks)
        out.copy_(torch.from_numpy(np.ascontiguousarray(img).reshape(-1)))


def _worker(rank, world, port, W, H, tile, result_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import distributed_raytracer_b200 as rt
        from distributed_raytracer_b200 import parallel
        from oracle import scene_loader as sl
        full = sl.environment_from_file(EXAMPLE_SCENE) if rank == 0 else None
        sc = parallel.broadcast_scene(full, src=0)
        if rank == 0:
            for k in ("verts", "normals", "faces", "materials"):
                assert np.array_equal(sc["meshes"][0][k], full["meshes"][0][k])
        # frame state (< 1 KB) is handed to every rank by its caller
        frame = [None]
        if rank == 0:
            frame[0] = (full["camera"], full["lights"])
        dist.broadcast_object_list(frame, src=0)
        cam, lights = frame[0]
        r = parallel.TileShardedRenderer(sc, W, H, precision=rt.FP64_EXACT, tile=tile, backend_factory=EmulTileBackend)
        assert r.n_ranks == world and r.rank == rank
        fb = r.render_frame(cam, lights)
        fb2 = r.render_frame(cam, lights)          # a second frame re-uses the buffers
        if rank == 0:
            assert torch.equal(fb, fb2)
            np.save(result_path, fb.numpy().reshape(H, W, 3))
        else:
            assert fb is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H,tile", [(96, 72, (32, 16)), (100, 70, (16, 8))])
def test_tile_sharded_frame_world2_gloo(tmp_path, built, W, H, tile):
    from oracle import oracle as orc
    from oracle import scene_loader as sl
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), W, H, tile, out), nprocs=2, join=True)
    img = np.load(out)
    sc = sl.environment_from_file(EXAMPLE_SCENE)
    rgb, _, _, _ = orc.Scene(sc).render_region(W, H, layout=orc.L_ROWMAJOR)
    assert np.array_equal(img, rgb.reshape(H, W, 3))
