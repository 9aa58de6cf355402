"""Multi-GPU screen-tile sharding: one process per GPU, torch.distributed (NCCL over NVLink) for plumbing.

The reference shards a frame by screen rectangles with no inter-worker exchange
(/root/reference/master/main.go:54-91 partition(), :166-176 result concatenation).  Here the frame is
cut into small tiles dealt round-robin to the ranks (only ~10 % of the pixels, in the middle of the
screen, carry all the shading + shadow work, so big rectangles would not balance):

  once per scene : rank 0 broadcasts the flattened SoA scene arrays; every rank builds the same
                   (deterministic) LBVH locally                         [dist.broadcast]
  per frame      : every rank gets the <1 KB camera/lights state from its caller (kernel parameters) and
                   makes ONE library call; no collective is on the data path:
      mode "peer"  rt_render_push_tiles: tiles rendered into a packed local buffer, stored into rank 0's
                   device framebuffer over NVLink (CUDA IPC), frame handed over through arrive/consumed
                   flags in rank 0's memory                              [device-resident frames]
      mode "host"  rt_render_stripes_host: stripes contiguous in the output layout rendered in place, one
                   strided copy-engine transfer per rank into a pinned host-shared framebuffer over the
                   rank's own PCIe link, flags in the same host memory   [frames in host memory: e2e]
      mode "peer_nccl" / "gather": the NCCL baselines kept for comparison (4-byte all_reduce barrier;
                   all_gather of packed tiles + rt_untile_device on rank 0).

The render backend is injectable so that the host-side logic (tile maps, gather, untile) can be
exercised on CPU with gloo in tests; the default backend is the GPU one and has no CPU fallback.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .gpu_scene import DeviceBuffer, FrameState, GpuScene, HostSharedBuffer, tiles_per_rank

import os as _os
_NO_BARRIER = bool(int(_os.environ.get("RT_B200_NO_BARRIER", "0")))   # timing experiments only: frames are then unsynchronised
_MESH_KEYS = (("verts", np.float64, 3), ("normals", np.float64, 3), ("faces", np.uint32, 7), ("materials", np.float64, 10))


def broadcast_scene(scene, src=0, group=None, device="cpu"):
    """Broadcasts the flattened scene arrays from `src`; returns the scene dict on every rank."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta[0] = {"n_meshes": len(scene["meshes"]),
                   "shapes": [[int(np.asarray(m[k]).reshape(-1, w).shape[0]) for k, _, w in _MESH_KEYS] for m in scene["meshes"]],
                   "n_obj": int(np.asarray(scene["obj_mesh"]).shape[0])}
    dist.broadcast_object_list(meta, src=src, group=group)
    meta = meta[0]
    out = {"meshes": []}

    def bcast(arr, dtype, shape):
        # uint32 / float64 payloads travel as raw bytes
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if rank == src:
            host = np.ascontiguousarray(arr, dtype=dtype).reshape(shape)
            t = torch.from_numpy(host.view(np.uint8).reshape(-1).copy())
        else:
            t = torch.empty(nbytes, dtype=torch.uint8)
        if nbytes == 0:
            return np.zeros(shape, dtype=dtype)
        t = t.to(device)
        dist.broadcast(t, src=src, group=group)
        return t.cpu().numpy().view(dtype).reshape(shape).copy()

    for k in range(meta["n_meshes"]):
        m = {}
        for (key, dtype, width), rows in zip(_MESH_KEYS, meta["shapes"][k]):
            m[key] = bcast(scene["meshes"][k][key] if rank == src else None, dtype, (rows, width))
        out["meshes"].append(m)
    n = meta["n_obj"]
    out["obj_pos"] = bcast(scene["obj_pos"] if rank == src else None, np.float64, (n, 3))
    out["obj_mesh"] = bcast(scene["obj_mesh"] if rank == src else None, np.uint32, (n,))
    return out


class GpuTileBackend:
    """Default backend: librt_b200.so on this rank's GPU (fails loudly without CUDA).

    `replicas` resident copies of the scene, each with its own scratch buffers and its own CUDA stream: consecutive frames
    go to consecutive replicas, so (a) up to `replicas` frames are in flight and the latency-bound tail of one frame's
    kernels (a few pathological rays with ~150 dependent traversal steps) overlaps the head of the next frame instead of
    idling the GPU, and (b) a frame's inputs were last touched `replicas` frames earlier and are no longer in L2, which is
    how the pipelined loop satisfies "inputs larger than L2" without a serialising flush."""

    pipelined = True

    def __init__(self, scene, device_index, precision, replicas=1, shared_scene=False):
        self.device = torch.device("cuda", device_index)
        n = max(1, replicas)
        # shared_scene: ONE resident scene, the frames in flight use its lanes 1..n (how a registered worker serves concurrent
        # orders); otherwise n resident copies (benchmarking: a frame's inputs were last touched n frames earlier)
        self.shared = bool(shared_scene)
        if self.shared:
            old = _os.environ.get("RT_B200_LANES")
            _os.environ["RT_B200_LANES"] = str(max(n + 1, int(old or 0)))       # read by rt_scene_create
            try:
                one = GpuScene(scene, device=device_index, precision=precision)
            finally:
                if old is None:
                    del _os.environ["RT_B200_LANES"]
                else:
                    _os.environ["RT_B200_LANES"] = old
            if one.lane_count() < n + 1:
                raise RuntimeError("scene has %d lanes, %d frames in flight requested (RT_B200_LANES)" % (one.lane_count(), n))
            self.scenes = [one] * n
            self.lane_flags = [_lib.LANE(k + 1) for k in range(n)]
            self.lanes = [k + 1 for k in range(n)]
        else:
            self.scenes = [GpuScene(scene, device=device_index, precision=precision) for _ in range(n)]
            self.lane_flags = [0] * n
            self.lanes = [0] * n
        self.streams = [torch.cuda.Stream(self.device) for _ in self.scenes]
        self.comm_stream = torch.cuda.Stream(self.device)

    @property
    def scene(self):
        return self.scenes[0]

    def n_slots(self):
        return len(self.scenes)

    def set_frame(self, slot, cam13, lights, obj_pos=None, tan_half_fov=float("nan")):
        self.set_frame_state(slot, FrameState(cam13, lights, obj_pos, tan_half_fov))

    def set_frame_state(self, slot, frame):
        self.scenes[slot].set_frame_state(frame, self.lanes[slot])     # host-only: the state travels as kernel parameters

    def alloc(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def render_region(self, slot, W, H, out, flags=0):
        self.scenes[slot].render_region_device(W, H, 0, 0, W, H, out.data_ptr(), self.streams[slot].cuda_stream,
                                               flags | self.lane_flags[slot])

    def render_tiles(self, slot, W, H, tw, th, rank, n_ranks, out, flags=0):
        self.scenes[slot].render_tiles_device(W, H, tw, th, rank, n_ranks, out.data_ptr(), self.streams[slot].cuda_stream,
                                              flags | self.lane_flags[slot])

    def render_tiles_peer(self, slot, W, H, tw, th, rank, n_ranks, frame_ptr, flags=0):
        self.scenes[slot].render_tiles_peer(W, H, tw, th, rank, n_ranks, frame_ptr, self.streams[slot].cuda_stream,
                                            flags | self.lane_flags[slot])

    def render_push_tiles(self, slot, frame, W, H, tw, th, rank, n_ranks, tiles, frame_ptr, flags=0, sync_slot_ptr=None, seq=0):
        self.scenes[slot].render_push_tiles(frame, W, H, tw, th, rank, n_ranks, tiles.data_ptr(), frame_ptr,
                                            self.streams[slot].cuda_stream, flags | self.lane_flags[slot], sync_slot_ptr, seq)

    def render_stripes_host(self, slot, frame, W, H, stripe, rank, n_ranks, local_fb, host_fb_ptr, flags=0, sync_slot_ptr=None, seq=0):
        self.scenes[slot].render_stripes_host(frame, W, H, stripe, rank, n_ranks, local_fb.data_ptr(), host_fb_ptr,
                                              self.streams[slot].cuda_stream, flags | self.lane_flags[slot], sync_slot_ptr, seq)

    def frame_complete(self, slot, sync_slot_ptr, seq, n_ranks, release):
        self.scenes[slot].frame_complete_device(sync_slot_ptr, seq, n_ranks, release, self.streams[slot].cuda_stream)

    def frame_release(self, slot, sync_slot_ptr, seq):
        self.scenes[slot].frame_release_device(sync_slot_ptr, seq, self.streams[slot].cuda_stream)

    def push_tiles(self, slot, W, H, tw, th, rank, n_ranks, tiles, frame_ptr, flags=0):
        self.scenes[slot].push_tiles_device(W, H, tw, th, rank, n_ranks, tiles.data_ptr(), frame_ptr, self.streams[slot].cuda_stream,
                                            flags | self.lane_flags[slot])

    def untile(self, W, H, tw, th, n_ranks, all_tiles, out, flags=0):
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.scenes[0].untile_device(W, H, tw, th, n_ranks, all_tiles.data_ptr(), out.data_ptr(), st, flags)


class TileShardedRenderer:
    """Renders full frames; with N > 1 ranks the screen tiles are dealt round-robin over the ranks of `group`.

    Frames are pipelined over the backend's slots (lanes of one scene, or scene copies; one stream each): frame k renders on
    slot k % S while earlier frames are still being handed over.  `render_frame` returns the framebuffer of
    the frame it enqueued (rank 0; None elsewhere); it is complete once `wait()` followed by a synchronize has returned.
    With one rank there is nothing to gather: the frame is rendered straight into the framebuffer (region mode).

    Lifetime of a device-resident frame: slot k = frame number % S is overwritten by the frame enqueued S frames later -- by any
    rank, as soon as rank 0 has released the slot.  Rank 0 releases it right after the frame is complete unless the frame has a
    consumer on the slot's stream (`to_host=True`: the copy into pinned host memory), in which case it is released after the
    consumer.  A caller that reads `frame[k]` by other means must finish before any rank enqueues frame k + S."""

    def __init__(self, scene, W, H, precision=_lib.FP32_FAST, tile=(32, 16), group=None, device_index=0, backend_factory=None,
                 layout_flags=_lib.LAYOUT_ROWMAJOR, replicas=1, mode="peer", shared_scene=False):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_ranks = dist.get_world_size(group) if dist.is_initialized() else 1
        self.W, self.H = int(W), int(H)
        self.tw, self.th = tile
        if self.tw % 8 or self.th % 4:
            raise ValueError("tile size must be a multiple of 8x4")
        self.layout_flags = layout_flags
        self.backend = backend_factory(scene, device_index, precision) if backend_factory else \
            GpuTileBackend(scene, device_index, precision, replicas, shared_scene)
        self.pipelined = bool(getattr(self.backend, "pipelined", False))
        S = self.backend.n_slots() if self.pipelined else 1
        self.S = S
        self.per_rank = tiles_per_rank(self.W, self.H, self.tw, self.th, self.n_ranks)
        self.tile_bytes = self.tw * self.th * 3
        multi = self.n_ranks > 1
        # mode "peer": every rank pushes its rendered tiles straight into rank 0's framebuffer over NVLink (CUDA IPC
        # mapping, wide coalesced stores) and the same kernel hands the frame over through flags in rank 0's memory
        # (arrive / consumed, release-acquire at system scope): no gather, no de-interleave, no collective, ONE library call
        # per frame and rank, and a rank only ever waits for rank 0's consumption of the slot, never for its siblings.
        # mode "peer_nccl": same data path, a 4-byte all_reduce per frame as the barrier (kept for comparison).
        # mode "gather": packed tile buffers + all_gather + de-interleave on rank 0 (the NCCL-only baseline).
        # mode "host": the framebuffers and their hand-off records live in HOST shared memory pinned by every rank, and the
        # screen is cut into stripes that are contiguous in the output layout (tile[0] pixel columns x full height for the
        # column-major layout, tile[1] pixel rows x full width for the row-major one).  Each rank renders its stripes in place
        # and moves them with ONE strided copy-engine transfer over its own PCIe link, so a frame reaches host memory at N x
        # the bandwidth of the single link a device->host copy from rank 0 has (frames are CPU tensors over the shared memory).
        # (Storing the tiles into host memory from the push kernel works too, rt_host_register, but reaches only ~11 GB/s
        # per link: profiles/r01_host_mode_kernel_stores_n2.jsonl.)
        if mode not in ("peer", "peer_nccl", "gather", "host"):
            raise ValueError("mode must be peer, peer_nccl, gather or host")
        self.mode = mode if (multi and self.pipelined) else "gather"
        self.peer_bufs = []
        self.ctrl = None
        self.slot_seq = [0] * S
        if self.mode in ("peer", "peer_nccl"):
            self._init_peer(device_index)
        elif self.mode == "host":
            self._init_host(device_index)
        if self.mode == "host":
            self.stripe = self.th if (layout_flags & _lib.LAYOUT_ROWMAJOR) else self.tw
            self.local = [self.backend.alloc(self.W * self.H * 3) for _ in range(S)]
        else:
            self.local = [self.backend.alloc(self.per_rank * self.tile_bytes) for _ in range(S)] if multi else []
        for t in self.local:
            t.zero_()
        self.gathered = [self.backend.alloc(self.n_ranks * self.per_rank * self.tile_bytes) for _ in range(S)] \
            if (multi and self.mode == "gather") else []
        if self.mode == "gather":
            self.frame = [self.backend.alloc(self.W * self.H * 3) if self.rank == 0 else None for _ in range(S)]
        self.seq = 0
        self.done_ev = [None] * S        # event: everything that read or wrote slot k's buffers is finished
        self._events = [(torch.cuda.Event(), torch.cuda.Event()) for _ in range(S)] if self.pipelined else []

    def _init_peer(self, device_index):
        nbytes = self.W * self.H * 3
        handles = [None]
        cbytes = self.S * _lib.SYNC_SLOT_BYTES
        if self.rank == 0:
            self.peer_bufs = [DeviceBuffer(device_index, nbytes) for _ in range(self.S)]
            self.ctrl = DeviceBuffer(device_index, cbytes)                 # frame hand-off records, one per slot
            handles = [[b.export() for b in self.peer_bufs] + [self.ctrl.export()]]
        dist.broadcast_object_list(handles, src=0, group=self.group)
        if self.rank != 0:
            self.peer_bufs = [DeviceBuffer(device_index, nbytes, ipc_handle=h) for h in handles[0][:-1]]
            self.ctrl = DeviceBuffer(device_index, cbytes, ipc_handle=handles[0][-1])
        self.frame = [torch.as_tensor(b, device=self.backend.device) if self.rank == 0 else None for b in self.peer_bufs]
        self.ctrl_view = torch.as_tensor(self.ctrl, device=self.backend.device).view(torch.int32).view(self.S, -1)
        if self.rank == 0:
            for f in self.frame:
                f.zero_()
            self.ctrl_view.zero_()
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.backend.device)
        torch.cuda.synchronize(self.backend.device)
        dist.barrier(group=self.group)

    def _init_host(self, device_index):
        nbytes = self.W * self.H * 3
        cbytes = self.S * _lib.SYNC_SLOT_BYTES
        names = [None]
        if self.rank == 0:
            try:
                st = _os.statvfs("/dev/shm")
                need = self.S * nbytes + cbytes
                if st.f_bavail * st.f_frsize < need + (64 << 20):
                    raise OSError("/dev/shm has %d MB free, %d MB needed" % (st.f_bavail * st.f_frsize >> 20, need >> 20))
                self.peer_bufs = [HostSharedBuffer(device_index, nbytes) for _ in range(self.S)]
                self.ctrl = HostSharedBuffer(device_index, cbytes)
                names = [[b.name for b in self.peer_bufs] + [self.ctrl.name]]
            except Exception as ex:                       # every rank must learn about it: the attach below is collective
                for b in self.peer_bufs:
                    b.close()
                self.peer_bufs = []
                names = ["ERROR: %r" % (ex,)]
        dist.broadcast_object_list(names, src=0, group=self.group)
        if isinstance(names[0], str):
            raise RuntimeError("host-shared framebuffers unavailable on rank 0: " + names[0])
        if self.rank != 0:
            self.peer_bufs = [HostSharedBuffer(device_index, name=n) for n in names[0][:-1]]
            self.ctrl = HostSharedBuffer(device_index, name=names[0][-1])
        self.frame = [torch.from_numpy(b.array[:nbytes]) if self.rank == 0 else None for b in self.peer_bufs]
        self.ctrl_view = torch.from_numpy(self.ctrl.array[:cbytes]).view(torch.int32).view(self.S, -1)
        torch.cuda.synchronize(self.backend.device)
        dist.barrier(group=self.group)

    def close(self):
        """Releases the shared framebuffers (host mode: unpins and, on rank 0, unlinks the shared-memory segments)."""
        if self.mode == "host":
            torch.cuda.synchronize(self.backend.device)
            self.frame, self.ctrl_view = [], None
            for b in self.peer_bufs + ([self.ctrl] if self.ctrl is not None else []):
                b.close()
            self.peer_bufs, self.ctrl = [], None

    def render_frame(self, cam13, lights=None, obj_pos=None, tan_half_fov=float("nan"), flags=0, to_host=False):
        """Enqueues one frame; asynchronous on CUDA.  `cam13` may be a prebuilt FrameState (then `lights` etc. are ignored).
        to_host: rank 0 also copies the finished framebuffer into its pinned host buffer `host_frame[slot]` (returned instead
        of the device tensor); valid after wait() + synchronize.  In mode "host" the frame is a host tensor anyway."""
        k = self.seq % self.S
        self.seq += 1
        if self.pipelined and not isinstance(cam13, FrameState) and (self.n_ranks == 1 or self.mode != "gather"):
            cam13 = FrameState(cam13, lights, obj_pos, tan_half_fov)
        if self.pipelined and isinstance(cam13, FrameState):
            return self._render_frame_fast(k, cam13, flags, to_host)
        if isinstance(cam13, FrameState):
            raise TypeError("FrameState needs the pipelined GPU backend")
        if not self.pipelined:
            # injected (test) backend: plain sequential path on the current stream
            self.backend.set_frame(cam13, lights, obj_pos, tan_half_fov)
            if self.n_ranks == 1:
                self.backend.render_tiles(self.W, self.H, self.tw, self.th, 0, 1, self._single_local(), flags)
                self.backend.untile(self.W, self.H, self.tw, self.th, 1, self._single_local(), self.frame[0], self.layout_flags)
                return self.frame[0]
            self.backend.render_tiles(self.W, self.H, self.tw, self.th, self.rank, self.n_ranks, self.local[0], flags)
            dist.all_gather_into_tensor(self.gathered[0], self.local[0], group=self.group)
            if self.rank == 0:
                self.backend.untile(self.W, self.H, self.tw, self.th, self.n_ranks, self.gathered[0], self.frame[0], self.layout_flags)
            return self.frame[0]
        be = self.backend
        st = be.streams[k]
        be.set_frame(k, cam13, lights, obj_pos, tan_half_fov)
        if self.done_ev[k] is not None:
            st.wait_event(self.done_ev[k])                 # slot k's previous frame (incl. its gather) must be finished
        # (one rank and the peer / host modes went through _render_frame_fast above: this is the pipelined NCCL gather)
        be.render_tiles(k, self.W, self.H, self.tw, self.th, self.rank, self.n_ranks, self.local[k], flags)
        ev = torch.cuda.Event()
        ev.record(st)
        comm = be.comm_stream
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(self.gathered[k], self.local[k], group=self.group)
            if self.rank == 0:
                be.untile(self.W, self.H, self.tw, self.th, self.n_ranks, self.gathered[k], self.frame[k], self.layout_flags)
            done = torch.cuda.Event()
            done.record(comm)
        self.done_ev[k] = done
        return self.frame[k]

    def _host_frames(self):
        if not hasattr(self, "host_frame"):
            self.host_frame = [torch.empty(self.W * self.H * 3, dtype=torch.uint8, pin_memory=True) if self.rank == 0 else None
                               for _ in range(self.S)]
        return self.host_frame

    def _render_frame_fast(self, k, frame, flags, to_host=False):
        be = self.backend
        st = be.streams[k]
        # Slot reuse.  Device-only frames: wait for the barrier of the frame that used this slot S frames ago.  With the
        # device->host copy on rank 0, the copy of frame f sits between the barriers of f and f+1 on rank 0's communication
        # stream, so "barrier of frame f+1 has completed" (a collective: it needs rank 0) tells EVERY rank that rank 0 has
        # finished reading slot k -- hence the wait on the next slot's event.
        if self.n_ranks > 1 and self.mode in ("peer", "host"):
            gate = None                                      # slot streams are in order; cross-rank reuse is gated by the consumed flag
        else:
            gate = self.done_ev[(k + 1) % self.S] if (to_host and self.S > 1) else self.done_ev[k]
        if gate is not None:
            st.wait_event(gate)
        if self.n_ranks == 1:
            be.set_frame_state(k, frame)
            be.render_region(k, self.W, self.H, self.frame[k], flags | self.layout_flags)
            ev = self._events[k][0]
            if to_host:
                with torch.cuda.stream(st):
                    self._host_frames()[k].copy_(self.frame[k], non_blocking=True)
            ev.record(st)
            self.done_ev[k] = ev
            return self._host_frames()[k] if to_host else self.frame[k]
        if self.mode == "gather":
            raise RuntimeError("FrameState fast path is implemented for the peer modes")
        ev, done = self._events[k]
        if self.mode == "host":
            # stripes rendered in place, one strided D2H transfer per rank, flags in the same pinned host shared memory
            self.slot_seq[k] += 1
            seq = self.slot_seq[k] & 0x7fffffff
            slot_ptr = self.ctrl.ptr + k * _lib.SYNC_SLOT_BYTES
            be.render_stripes_host(k, frame, self.W, self.H, self.stripe, self.rank, self.n_ranks, self.local[k], self.peer_bufs[k].ptr,
                                   flags | self.layout_flags, slot_ptr, seq)
            if self.rank == 0:
                be.frame_complete(k, slot_ptr, seq, self.n_ranks, release=True)
            done.record(st)
            self.done_ev[k] = done
            return self.frame[k]
        if self.mode == "peer":
            # ONE stream, no collective: render -> push (waits for "slot consumed", publishes "arrived") [-> rank 0: complete
            # (waits for every rank's arrival) -> consumer -> release]
            self.slot_seq[k] += 1
            seq = self.slot_seq[k] & 0x7fffffff
            slot_ptr = self.ctrl.ptr + k * _lib.SYNC_SLOT_BYTES
            be.render_push_tiles(k, frame, self.W, self.H, self.tw, self.th, self.rank, self.n_ranks, self.local[k], self.peer_bufs[k].ptr,
                                 flags | self.layout_flags, slot_ptr, seq)
            if self.rank == 0:
                be.frame_complete(k, slot_ptr, seq, self.n_ranks, release=not to_host)
                if to_host:
                    with torch.cuda.stream(st):
                        self._host_frames()[k].copy_(self.frame[k], non_blocking=True)
                    be.frame_release(k, slot_ptr, seq)
            done.record(st)
            self.done_ev[k] = done
            return self._host_frames()[k] if to_host else self.frame[k]
        be.render_push_tiles(k, frame, self.W, self.H, self.tw, self.th, self.rank, self.n_ranks, self.local[k], self.peer_bufs[k].ptr,
                             flags | self.layout_flags)
        ev.record(st)
        comm = be.comm_stream
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            if not _NO_BARRIER:
                dist.all_reduce(self.flag, group=self.group)
            if to_host and self.rank == 0:
                self._host_frames()[k].copy_(self.frame[k], non_blocking=True)
            done.record(comm)
        self.done_ev[k] = done
        if to_host:
            return self._host_frames()[k]
        return self.frame[k]

    def check_sync(self, where=""):
        """Raises if a device-side wait of the frame hand-off timed out (a rank died or fell > RT_B200_SYNC_TIMEOUT_MS behind)."""
        if self.ctrl is None:
            return
        words = self.ctrl_view.cpu().clone()
        err = words[:, 33]
        if int(err.abs().max()) != 0:
            raise RuntimeError("frame hand-off timed out%s: error words per slot = %s (0x100+r: rank r never arrived; 0x200+r: "
                               "rank r never saw the slot released); arrive[0..%d] / consumed per slot = %s; frames enqueued per slot here = %s"
                               % (where, [hex(int(e)) for e in err], self.n_ranks - 1,
                                  [([int(v) for v in w[:self.n_ranks]], int(w[32])) for w in words], self.slot_seq))

    def clear_sync_errors(self):
        """Rank 0 zeroes the error words of the hand-off records (callers put a barrier on both sides)."""
        if self.ctrl is not None and self.rank == 0:
            self.ctrl_view[:, 33] = 0
            if self.backend is not None and hasattr(self.backend, "device"):
                torch.cuda.synchronize(self.backend.device)

    def _single_local(self):
        if not self.local:
            self.local = [self.backend.alloc(self.per_rank * self.tile_bytes)]
        return self.local[0]

    def start_after(self, event):
        """Makes every slot stream (and the communication stream) wait for `event` (start of a timed region)."""
        if self.pipelined:
            for st in self.backend.streams:
                st.wait_event(event)
            self.backend.comm_stream.wait_event(event)

    def wait(self):
        """Makes the caller's current stream wait for every frame enqueued so far."""
        if self.pipelined:
            cur = torch.cuda.current_stream(self.backend.device)
            for e in self.done_ev:
                if e is not None:
                    cur.wait_event(e)
