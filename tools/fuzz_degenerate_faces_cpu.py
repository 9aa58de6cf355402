"""CPU-only stress of the zero-area-face corner (profiles/r01_fuzz_cpu_replay.txt): scenes snapped to a 0.25 / 0.5 grid so that about
half of the faces are degenerate; compares the CPU replay of the kernels (zero-area faces
met through box.go's predicate, rt_exact.cuh) with the oracle's R-tree mode (reference-shaped: no distance pruning); the
oracle's distance-pruning mode is counted beside it to show how often the corner matters.   python tools/fuzz_degenerate_faces_cpu.py <seed> <seconds>"""
import sys, time, pickle
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import random_soup_scene
from oracle import oracle as orc, scene_loader as sl
import emul_build as eb
import distributed_raytracer_b200 as rt
rng = np.random.default_rng(int(sys.argv[1]))
bad = 0; bad_cr = 0; t0=time.time(); n=0; ndeg=0
while time.time()-t0 < float(sys.argv[2]):
    kw = dict(n_tris=int(rng.integers(20, 300)), seed=int(rng.integers(0, 1<<30)), with_normals=bool(rng.integers(0,2)), n_obj=int(rng.integers(1,3)),
              n_lights=int(rng.integers(1,4)), extent=float(rng.choice([1.0,2.5])), half=float(rng.choice([0.15,0.6])))
    sc = random_soup_scene(**kw)
    pos = rng.normal(size=3)*2 + np.array([0,0,5.0]); d = -pos + rng.normal(size=3)*0.5
    if np.linalg.norm(np.cross(d,[0,1,0])) < 1e-3: continue
    sc["camera"] = sl.new_camera(pos, d, float(rng.uniform(0.5, 1.5)))
    g = float(rng.choice([2,4]))
    v = sc["meshes"][0]["verts"]; sc["meshes"][0]["verts"] = (np.round(v*g)/g).astype(np.float32).astype(np.float64)
    V = sc["meshes"][0]["verts"].reshape(-1,3,3)
    area = np.linalg.norm(np.cross(V[:,1]-V[:,0], V[:,2]-V[:,0]), axis=1)
    ndeg += int((area==0).sum())
    W,H = 32, 24
    S = orc.Scene(sc); S.build_rtrees()
    rgb_c, col_c, face_c, _ = S.render_region(W,H,mode=orc.CULLED)
    rgb_r, col_r, face_r, _ = S.render_region(W,H,mode=orc.RTREE)
    E = eb.EmulScene(sc, rt.FP64_EXACT, 4)
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H)
    same_cr = np.array_equal(face_c, face_r) and np.array_equal(col_c.view(np.uint64), col_r.view(np.uint64))
    same_ec = np.array_equal(face, face_c) and np.array_equal(col.view(np.uint64), col_c.view(np.uint64))
    same_er = np.array_equal(face, face_r) and np.array_equal(col.view(np.uint64), col_r.view(np.uint64))     # the claim: replay == reference-shaped walk
    n+=1
    bad_cr += (not same_cr)
    if not same_er:
        bad+=1
        npx = int(((col_c.view(np.uint64)!=col_r.view(np.uint64)).any(axis=1) | (face_c!=face_r)).sum())
        npx = int(((col.view(np.uint64)!=col_r.view(np.uint64)).any(axis=1) | (face!=face_r)).sum())
        print("DIFF emul==rtree", same_er, "culled==rtree", same_cr, "emul==culled", same_ec, "pixels", npx, kw, flush=True)
        pickle.dump((sc,W,H), open("/tmp/fuzz2_fail_%s_%d.pkl"%(sys.argv[1],bad),"wb"))
    E.close(); S.close()
print("scenes", n, "degenerate faces total", ndeg, "replay != R-tree walk:", bad, "| distance-pruning oracle mode != R-tree walk:", bad_cr)
