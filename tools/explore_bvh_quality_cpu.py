"""CPU-only experiment (tests/emul replay of the device code): node visits and triangle tests per frame of C3 at 960x540 for
the Karras LBVH, an SAH sweep along the Morton order, and a top-down binned SAH tree, all fed through the same refit /
4-wide emission / traversal code.   python tools/explore_bvh_quality_cpu.py karras|sweep|sah"""
import os, sys, time, subprocess, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
mode = sys.argv[1]
if mode != "karras": os.environ["RT_B200_BUILDER"] = mode
import numpy as np
import bench
import distributed_raytracer_b200 as rt
import emul_build
sc = bench.build_workload()
t=time.time(); E = emul_build.EmulScene(sc, rt.FP32_FAST); tb=time.time()-t
rc, nodes, depth, leaves = E.bvh_check()
W,H = 960,540
t=time.time()
rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H, tan_half_fov=sc["tan_half_fov"])
tr=time.time()-t
hits=int((face>=0).sum())
print(json.dumps({"builder":mode,"build_s":tb,"bvh_rc":rc,"nodes":nodes,"depth":depth,"leaves":leaves,"render_s":tr,"hits":hits,
  "node_visits":int(cnt[0]),"tri_tests":int(cnt[1]),"per_ray_nodes":float(cnt[0])/(W*H+2*hits),"per_ray_tris":float(cnt[1])/(W*H+2*hits),
  "checksum":int(rgb.astype(np.uint64).sum())}))
