"""CPU-only fuzzer: random OBJ/MTL files (triangles and quads, v / v/vt / v//vn / v/vt/vn corners, negative indices, duplicate
vertices, zero normals, o/g/s/usemtl lines, materials outside [0,1]) loaded by the C++ mirror of the reference loader
(host/state.cpp) and by the independent Python one (oracle/scene_loader.py): every array must agree bit for bit.
   python tools/fuzz_loader_cpu.py <seed> <seconds>      Round 1: 3 x 80 s, 26 931 files, no disagreement."""
import sys, os, json, tempfile, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import scene_loader as sl
from distributed_raytracer_b200 import host
rng = np.random.default_rng(int(sys.argv[1]))
tmp = tempfile.mkdtemp()
t0=time.time(); n=0; bad=0
def fmt(x):
    s = rng.integers(0,4)
    return ["%.6f"%x, "%g"%x, "%.3e"%x, "%d"%round(x)][s]
while time.time()-t0 < float(sys.argv[2]):
    nv = int(rng.integers(3, 30)); nn = int(rng.integers(0, 20))
    lines = ["# fuzz", "mtllib m.mtl"]
    names = ["A","B","C"]
    mtl = []
    for nm in names[:int(rng.integers(0,4))]:
        mtl += ["newmtl %s"%nm, "Ns %s"%fmt(float(rng.integers(0,20))), "Ka %s %s %s"%tuple(fmt(v) for v in rng.uniform(-0.2,1.3,3)),
                "Kd %s %s %s"%tuple(fmt(v) for v in rng.uniform(0,1,3)), "Ks %s %s %s"%tuple(fmt(v) for v in rng.uniform(0,1,3)), "d 1", "illum 2", ""]
    # duplicate some vertices on purpose (value-keyed dedup, mesh.go:146-203)
    V = rng.integers(-3,4,size=(nv,3))/2.0 if rng.integers(0,2) else rng.normal(size=(nv,3))
    for v in V: lines.append("v %s %s %s"%tuple(fmt(c) for c in v))
    for k in range(nn):
        v = rng.normal(size=3); lines.append("vn %s %s %s"%tuple(fmt(c) for c in v))
    lines.append("vt 0.5 0.5")
    nf = int(rng.integers(1, 25))
    for f in range(nf):
        r = rng.integers(0,8)
        if r==0: lines.append("usemtl %s"%names[int(rng.integers(0,3))])
        if r==1: lines.append("g grp%d"%f)
        if r==2: lines.append("o obj%d"%f)
        if r==3: lines.append("s off")
        k = 3 if rng.integers(0,3) else 4
        use_n = nn>0 and rng.integers(0,2)
        neg = rng.integers(0,4)==0
        toks=[]
        for c in range(k):
            vi = int(rng.integers(1,nv+1)); vi_s = str(vi - nv - 1) if neg else str(vi)
            if use_n:
                ni = int(rng.integers(1,nn+1)); ni_s = str(ni - nn - 1) if neg else str(ni)
                toks.append("%s/%s/%s"%(vi_s, "1" if rng.integers(0,2) else "", ni_s))
            else:
                toks.append(vi_s if rng.integers(0,2) else vi_s+"/1")
        lines.append("f "+" ".join(toks))
    open(os.path.join(tmp,"m.obj"),"w").write("\n".join(lines)+"\n")
    open(os.path.join(tmp,"m.mtl"),"w").write("\n".join(mtl)+"\n")
    scene = {"objs":[{"model":"m.obj","pos":{"x":0.5,"y":-1.0,"z":2.0}}], "lights":[{"pos":{"x":0,"y":0,"z":10},"col":{"r":255,"g":128,"b":0}}],
             "cam":{"pos":{"x":0,"y":0,"z":5},"dir":{"x":0.1,"y":0,"z":-1},"fov":1.0}}
    open(os.path.join(tmp,"scene.json"),"w").write(json.dumps(scene))
    n+=1
    try:
        a = sl.environment_from_file(os.path.join(tmp,"scene.json")); ea=None
    except Exception as e: a=None; ea=e
    try:
        env = host.Environment(os.path.join(tmp,"scene.json")); b = env.scene_arrays(); env.close(); eb_=None
    except Exception as e: b=None; eb_=e
    ok = (a is None) == (b is None)
    if ok and a is not None:
        ma, mb = a["meshes"][0], b["meshes"][0]
        for key in ("verts","normals","faces","materials"):
            xa, xb = np.asarray(ma[key]), np.asarray(mb[key])
            if xa.shape != xb.shape or not np.array_equal(xa.view(np.uint8) if xa.dtype!=np.uint32 else xa, xb.view(np.uint8) if xb.dtype!=np.uint32 else xb):
                ok=False; why=key; break
        if ok and not (np.array_equal(a["camera"], b["camera"]) and np.array_equal(a["lights"], b["lights"])): ok=False; why="cam/lights"
    elif not ok: why="error mismatch: py=%r cpp=%r"%(ea, eb_)
    if not ok:
        bad+=1; print("MISMATCH", why); 
        import shutil; shutil.copytree(tmp, "/tmp/loader_fail_%d"%bad, dirs_exist_ok=True)
        if bad>3: break
print("files", n, "bad", bad)
