"""CPU-only model of K4's SIMD efficiency.  The CPU replay of the device code (tests/emul) logs, for every shadow ray of a C3
frame, the order in which it visits nodes and leaves; this script groups the rays into warps exactly as k_shade does (32
consecutive hits, one light at a time) and counts warp-level issue slots for different loop organisations:

  A  the product's loop: every iteration each lane handles its next item; the warp pays the node body if any lane is at a node
     and the leaf body (x the largest triangle count among those lanes) if any lane is at a leaf
  B  while-while: all lanes walk nodes until each has reached a leaf (or finished), then all pending leaves are processed
  C  A with the two lights' rays of a hit side by side (16 hits x 2 lights per warp instead of 32 hits, light after light)
  D  lower bound: perfect packing (thread-level work / 32)
  E  A with dynamic refill: a lane that finishes takes the next ray at once (removes the ray-length tail, not the node/leaf split)
  G_T  leaf postponement with a warp vote: an iteration runs the node body OR the leaf body; leaves run once >= T lanes wait
       at one (T = 32 is B)

Costs in issue slots: node body 100, leaf entry 15 + 60 per triangle, loop overhead 10 per iteration (counted once per warp
iteration).  The absolute numbers are a model; the RATIOS between organisations are what it is for.

   python tools/simulate_warp_divergence_cpu.py [width height [max_leaf]]      (default 960 540 4)
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

C_NODE, C_LEAF, C_TRI, C_LOOP = 100, 15, 60, 10


def ray_sequences(W, H, max_leaf=4):
    import bench
    import distributed_raytracer_b200 as rt
    import emul_build
    sc = bench.build_workload()
    E = emul_build.EmulScene(sc, rt.FP32_FAST, max_leaf)
    L = emul_build.lib()
    L.emul_shade_events.restype = C.c_longlong
    L.emul_shade_events.argtypes = [C.c_void_p, C.POINTER(rt._lib.Camera), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.c_ulonglong, C.POINTER(C.c_uint32)]
    cam = rt.make_camera(sc["camera"], sc["tan_half_fov"])
    lights = np.ascontiguousarray(sc["lights"], dtype=np.float64)
    buf = np.zeros(256 << 20, dtype=np.uint8)
    nh = C.c_uint32(0)
    n = L.emul_shade_events(E.h, C.byref(cam), lights.ctypes.data, lights.shape[0], W, H, buf.ctypes.data, buf.size, C.byref(nh))
    assert 0 < n <= buf.size
    ev = buf[:n]
    hits = []                      # per hit: list (per light) of event lists
    cur, ray, inside = [], None, False
    for b in ev.tolist():
        if b == 255:
            hits.append(cur); cur = []
        elif b == 254:
            if not inside:
                ray = []; inside = True
            else:
                cur.append(ray); inside = False
        else:
            ray.append(b)
    E.close()
    return hits, lights.shape[0]


def cost_A(lanes):
    """lanes: list of event lists (one per lane)."""
    pos = [0] * len(lanes)
    total = 0
    while True:
        at_node = False
        max_tris = 0
        active = 0
        for k, seq in enumerate(lanes):
            if pos[k] < len(seq):
                active += 1
                e = seq[pos[k]]
                if e == 0:
                    at_node = True
                else:
                    max_tris = max(max_tris, e)
                pos[k] += 1
        if not active:
            return total
        total += C_LOOP + (C_NODE if at_node else 0) + ((C_LEAF + C_TRI * max_tris) if max_tris else 0)


def cost_B(lanes):
    pos = [0] * len(lanes)
    total = 0
    while True:
        # phase 1: nodes until every live lane is at a leaf or done
        steps = 0
        for k, seq in enumerate(lanes):
            s = 0
            while pos[k] < len(seq) and seq[pos[k]] == 0:
                pos[k] += 1; s += 1
            steps = max(steps, s)
        total += steps * (C_NODE + C_LOOP)
        # phase 2: pending leaves
        max_tris = 0
        for k, seq in enumerate(lanes):
            if pos[k] < len(seq):
                max_tris = max(max_tris, seq[pos[k]]); pos[k] += 1
        if max_tris:
            total += C_LOOP + C_LEAF + C_TRI * max_tris
        elif steps == 0:
            return total


def cost_G(lanes, T, vote_cost=6):
    """leaf postponement with a threshold: an iteration runs EITHER the node body (all lanes that are at a node advance) OR the
    leaf body (all lanes that are at a leaf advance); the leaf body runs when at least T lanes wait at a leaf or no lane is at a
    node.  T = 32 is the while-while loop, T = 1 degenerates to "leaves first"."""
    pos = [0] * len(lanes)
    total = 0
    while True:
        n_node = n_leaf = max_tris = 0
        for k, seq in enumerate(lanes):
            if pos[k] < len(seq):
                if seq[pos[k]] == 0:
                    n_node += 1
                else:
                    n_leaf += 1; max_tris = max(max_tris, seq[pos[k]])
        if n_node == 0 and n_leaf == 0:
            return total
        total += C_LOOP + vote_cost
        if n_leaf >= T or n_node == 0:
            total += C_LEAF + C_TRI * max_tris
            for k, seq in enumerate(lanes):
                if pos[k] < len(seq) and seq[pos[k]] != 0:
                    pos[k] += 1
        else:
            total += C_NODE
            for k, seq in enumerate(lanes):
                if pos[k] < len(seq) and seq[pos[k]] == 0:
                    pos[k] += 1


def cost_E(queue, n_lanes=32):
    """the product's loop with dynamic refill: a lane that finishes its ray takes the next one from `queue` at once."""
    q = list(queue)
    nxt = 0
    lanes = []
    for _ in range(n_lanes):
        if nxt < len(q):
            lanes.append([q[nxt], 0]); nxt += 1
    total = 0
    while True:
        at_node = False
        max_tris = 0
        active = 0
        for ln in lanes:
            while ln[0] is not None and ln[1] >= len(ln[0]):     # finished (or empty) ray: refill
                if nxt < len(q):
                    ln[0], ln[1] = q[nxt], 0; nxt += 1
                else:
                    ln[0] = None
            if ln[0] is None:
                continue
            active += 1
            e = ln[0][ln[1]]
            if e == 0:
                at_node = True
            else:
                max_tris = max(max_tris, e)
            ln[1] += 1
        if not active:
            return total
        total += C_LOOP + (C_NODE if at_node else 0) + ((C_LEAF + C_TRI * max_tris) if max_tris else 0)


def thread_work(seq):
    return sum((C_NODE + C_LOOP) if e == 0 else (C_LOOP + C_LEAF + C_TRI * e) for e in seq)


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
    max_leaf = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    hits, n_lights = ray_sequences(W, H, max_leaf)
    tot = {"A": 0, "B": 0, "C": 0, "D": 0.0, "E_refill_1024": 0}
    THRESH = (2, 4, 6, 8, 12, 16, 24, 32)
    for T in THRESH:
        tot["G_T%d" % T] = 0
    nodes = leaves = 0
    for w0 in range(0, len(hits), 32):
        warp = hits[w0:w0 + 32]
        for l in range(n_lights):
            lanes = [h[l] for h in warp]
            tot["A"] += cost_A(lanes)
            tot["B"] += cost_B(lanes)
            for T in THRESH:
                tot["G_T%d" % T] += cost_G(lanes, T)
            tot["D"] += sum(thread_work(s) for s in lanes) / 32.0
            nodes += sum(s.count(0) for s in lanes); leaves += sum(len(s) - s.count(0) for s in lanes)
    for w0 in range(0, len(hits), 16):                      # C: 16 hits x all lights in one warp
        warp = hits[w0:w0 + 16]
        lanes = [h[l] for h in warp for l in range(n_lights)]
        for g in range(0, len(lanes), 32):
            tot["C"] += cost_A(lanes[g:g + 32])
    for w0 in range(0, len(hits), 512):                     # E: one warp works through 512 hits x all lights with refill
        q = [h[l] for h in hits[w0:w0 + 512] for l in range(n_lights)]
        tot["E_refill_1024"] += cost_E(q)
    n_rays = len(hits) * n_lights
    out = {"screen": [W, H], "max_leaf": max_leaf, "hits": len(hits), "shadow_rays": n_rays, "nodes_per_ray": nodes / n_rays, "leaves_per_ray": leaves / n_rays,
           "issue_slots_per_ray": {k: v / n_rays for k, v in tot.items()},
           "simd_efficiency": {k: tot["D"] / v for k, v in tot.items()},
           "relative_to_product_loop": {k: v / tot["A"] for k, v in tot.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
