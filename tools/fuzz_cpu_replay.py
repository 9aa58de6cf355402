"""CPU-only fuzzer: random multi-object scenes (some snapped to a coarse grid to provoke ties and degenerate faces), random
cameras, tiny screens; the CPU replay of the device code (tests/emul, FP64_EXACT) must equal the oracle bit for bit, the
oracle's brute-force and culled modes must agree, FP32_FAST must stay within tolerance.   python tools/fuzz_cpu_replay.py <seed> <seconds>
Round 1: 4 x 100 s, 52 512 scenes, one mismatch, explained in profiles/r01_fuzz_cpu_replay.txt."""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import random_soup_scene
from oracle import oracle as orc, scene_loader as sl
import emul_build as eb
import distributed_raytracer_b200 as rt
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
bad = 0; t0=time.time(); n=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 60:
    kw = dict(n_tris=int(rng.integers(1, 400)), seed=int(rng.integers(0, 1<<30)), with_normals=bool(rng.integers(0,2)), n_obj=int(rng.integers(1,4)),
              n_lights=int(rng.integers(0,5)), extent=float(rng.choice([0.3,1.0,2.5])), half=float(rng.choice([0.02,0.15,0.6])))
    sc = random_soup_scene(**kw)
    # random camera
    pos = rng.normal(size=3)*3 + np.array([0,0,4.0]); d = -pos + rng.normal(size=3)*0.5
    if np.linalg.norm(np.cross(d,[0,1,0])) < 1e-3: continue
    sc["camera"] = sl.new_camera(pos, d, float(rng.uniform(0.3, 2.2)))
    # snap some vertices to a coarse grid to provoke ties / degenerate triangles
    if rng.integers(0,3)==0:
        v = sc["meshes"][0]["verts"]; sc["meshes"][0]["verts"] = (np.round(v*4)/4).astype(np.float32).astype(np.float64)
    W,H = int(rng.integers(1,40)), int(rng.integers(1,30))
    ml = int(rng.choice([1,2,4,8]))
    S = orc.Scene(sc)
    rgb_o, col_o, face_o, st = S.render_region(W,H,mode=orc.CULLED)
    rgb_b, col_b, face_b, _ = S.render_region(W,H,mode=orc.BRUTE)
    E = eb.EmulScene(sc, rt.FP64_EXACT, ml)
    rc = E.bvh_check()[0]
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H)
    ok = rc==0 and np.array_equal(face, face_o) and np.array_equal(col.view(np.uint64), col_o.view(np.uint64)) and np.array_equal(rgb, rgb_o) \
         and np.array_equal(face_b, face_o) and np.array_equal(col_b.view(np.uint64), col_o.view(np.uint64))
    Ef = eb.EmulScene(sc, rt.FP32_FAST, ml)
    rgbf = Ef.render(sc["camera"], sc["lights"], W, H)[0]
    within = (np.abs(rgbf.astype(int)-rgb_o.astype(int)).max(axis=1)<=1).mean()
    n+=1
    if not ok or (within < 0.9 and W*H>200):
        bad+=1; import pickle; pickle.dump((sc,W,H,ml), open("/tmp/fuzz_fail_%d.pkl"%bad,"wb")); print("MISMATCH", kw, W, H, ml, rc, within, flush=True)
    E.close(); Ef.close(); S.close()
print("scenes", n, "bad", bad)
