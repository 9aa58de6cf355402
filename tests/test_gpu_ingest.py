"""GPU (-m gpu): SURVEY s8 row f3 -- registration straight from OBJ text, parsed on the GPU (csrc/obj_ingest.cu,
rt_scene_create_obj, host.gpu_scene_from_file).  The ingested scene must render exactly what the scene loaded by the independent
Python loader (oracle/scene_loader.py: gwob semantics, value-keyed de-duplication as in mesh.go:146-203) renders."""
import json
import os
import time

import numpy as np
import pytest

from conftest import EXAMPLE_SCENE, GOLDEN
from oracle import oracle as orc
from oracle import scene_loader as sl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built):
    import distributed_raytracer_b200 as rt
    if rt._lib.load().rt_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return rt


def test_example_scene_from_obj_text_matches_the_independent_golden(rt):
    from distributed_raytracer_b200 import host
    g, cam13, tan_half, lights, op = host.gpu_scene_from_file(EXAMPLE_SCENE, precision=rt.FP64_EXACT)
    info = g.bvh_info()
    assert info["n_faces"] == 968 and info["n_verts_ingested"] == 507 and info["n_normals_ingested"] == 507 and info["ms_ingest"] > 0
    rgb, col, face, obj, lit = g.render_region_debug(320, 240)
    g.close()
    gold = np.load(os.path.join(GOLDEN, "c1_320x240_pyref.npz"))
    hit = gold["hit_index"]
    assert np.array_equal(face, gold["face"].astype(np.int32))                   # face index = emission order of the OBJ file
    assert np.array_equal(col[hit].view(np.uint64), gold["colour_bits"])
    assert np.array_equal(rgb[hit], gold["rgb"]) and np.array_equal(lit[hit], gold["lit"].astype(np.uint32))


TRICKY_OBJ = """# every syntax the parser must take: CRLF, comments, exponents, signs, quads, negative indices, v / v/t / v//n / v/t/n
mtllib tricky.mtl\r
o thing
v  -1.0 -1.0 0.25
v 1.0e0 -1 0.25\r
v +1 1. .25
v -1 1 2.5e-1
vt 0 0
vt 1 1
vn 0 0 2
vn 0.0 1e-3 1.0
usemtl red
f 1/1/1 2/2/1 3/1/2 4/2/2
v 0.3 0.2 -0.5
v 1.9 0.1 -0.75
v 1.2 1.6 -0.625
usemtl blue
f -3//1 -2//2 -1//1
usemtl red
f 5//2 7//2 6//1
   # indented comment
f 1//1 3//2 2//1
"""
TRICKY_MTL = """newmtl red
Ns 10
Ka 0.1 0.0 0.0
Kd 0.8 0.1 0.1
Ks 0.5 0.5 0.5
newmtl blue
Ns 3
Ka 0.0 0.0 0.2
Kd 0.1 0.2 0.9
Ks 0.25 0.25 0.25
"""


def write_scene(tmp_path, obj_text, mtl_text, name="tricky"):
    (tmp_path / (name + ".obj")).write_bytes(obj_text.encode())
    (tmp_path / (name + ".mtl")).write_text(mtl_text)
    scene = {"objs": [{"model": name + ".obj", "pos": {"x": 0.0, "y": 0.0, "z": -1.0}}, {"model": name + ".obj", "pos": {"x": -2.5, "y": 0.5, "z": -2.0}}],
             "lights": [{"pos": {"x": 0.0, "y": 0.0, "z": 10.0}, "col": {"r": 255, "g": 255, "b": 255}},
                        {"pos": {"x": 5.0, "y": 8.0, "z": 6.0}, "col": {"r": 0, "g": 128, "b": 255}}],
             "cam": {"pos": {"x": 0.2, "y": 0.3, "z": 4.0}, "dir": {"x": 0.0, "y": -0.05, "z": -1.0}, "fov": 1.0}}
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(scene))
    return str(p)


def test_tricky_obj_syntax_matches_the_python_loader(rt, tmp_path):
    from distributed_raytracer_b200 import host
    path = write_scene(tmp_path, TRICKY_OBJ, TRICKY_MTL)
    sc = sl.environment_from_file(path)
    assert sc["meshes"][0]["faces"].shape[0] == 5 and sc["meshes"][0]["materials"].shape[0] == 2
    S = orc.Scene(sc)
    rgb_o, col_o, face_o, st = S.render_region(200, 150, mode=orc.CULLED)
    S.close()
    assert st["hit_pixels"] > 1000
    g, cam13, tan_half, lights, op = host.gpu_scene_from_file(path, precision=rt.FP64_EXACT)
    assert np.array_equal(cam13.view(np.uint64), np.asarray(sc["camera"], dtype=np.float64).view(np.uint64))
    assert np.array_equal(lights.view(np.uint64), sc["lights"].view(np.uint64)) and np.array_equal(op, sc["obj_pos"])
    rgb, col, face, obj, lit = g.render_region_debug(200, 150)
    info = g.bvh_info()
    g.close()
    assert info["n_faces"] == 5 and info["n_verts_ingested"] == 7 and info["n_normals_ingested"] == 2
    assert np.array_equal(face, face_o) and np.array_equal(col.view(np.uint64), col_o.view(np.uint64)) and np.array_equal(rgb, rgb_o)


@pytest.mark.parametrize("bad,needle", [("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nv 2 2 0\nf 1 2 3 4 5\n", "vertex count"),
                                        ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n", "out of range"),
                                        ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1 2 3\n", "with and without normal")])
def test_obj_ingest_rejections(rt, tmp_path, bad, needle):
    """More than four corners (gwob: triangles and quads only), an index out of range (the reference panics), and -- a documented
    limit of the GPU ingest -- a file that mixes faces with and without normal indices."""
    from distributed_raytracer_b200 import host
    path = write_scene(tmp_path, bad, "", name="bad")
    with pytest.raises(RuntimeError) as ei:
        host.gpu_scene_from_file(path)
    assert needle in str(ei.value)


def test_million_triangle_obj_text(rt, tmp_path, example_scene):
    """Config 3 as a 100 MB OBJ file: ingested on the GPU, it must render the frame the array registration renders; the time of
    the ingest is recorded next to the host loader's (the C++ mirror of MeshFromFile with its de-duplication maps)."""
    import torch
    from distributed_raytracer_b200 import host, workloads
    m = workloads.subdivide(example_scene["meshes"][0], 5)
    v, n, f = m["verts"], m["normals"], m["faces"]
    t0 = time.perf_counter()
    lines = ["mtllib big.mtl", "usemtl None"]
    lines += ["v %.9g %.9g %.9g" % tuple(r) for r in v]
    # vn lines carry UN-normalised float32 values in a real file; write the unit normals rounded to float32 and let both loaders
    # normalise them again (the comparison below is loader against loader on the same text)
    n32 = n.astype(np.float32)
    lines += ["vn %.9g %.9g %.9g" % tuple(r) for r in n32]
    lines += ["f %d//%d %d//%d %d//%d" % (a + 1, d + 1, b + 1, e + 1, c + 1, g + 1) for a, b, c, d, e, g in f[:, :6].tolist()]
    text = "\n".join(lines) + "\n"
    t_write = time.perf_counter() - t0
    path = write_scene(tmp_path, text, open(os.path.join(GOLDEN, "example", "suzanne.mtl")).read(), name="big")
    # only the first object of write_scene's two is kept: one 991 232-face object at the example position
    js = json.load(open(path))
    js["objs"] = [{"model": "big.obj", "pos": {"x": 1.0, "y": 1.0, "z": -1.0}}]
    js["lights"] = js["lights"][:2]
    js["cam"] = {"pos": {"x": 1.0, "y": 1.0, "z": 5.0}, "dir": {"x": 0.0, "y": 0.0, "z": -1.0}, "fov": 1.04719755}
    open(path, "w").write(json.dumps(js))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_bytes, n_mat, n_use = host.obj_text_probe(str(tmp_path / "big.obj"))
    t_probe = time.perf_counter() - t0
    t0 = time.perf_counter()
    g, cam13, tan_half, lights, op = host.gpu_scene_from_file(path, precision=rt.FP32_FAST)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    info = g.bvh_info()
    assert info["n_faces"] == 991232 and info["n_verts_ingested"] == v.shape[0] and info["n_normals_ingested"] == n.shape[0]
    W, H = 1920, 1080
    a = g.render_region(W, H)
    g.close()
    # the host loader (C++ mirror of MeshFromFile: tokenizer + value-keyed maps) on the same file, then the array registration
    t0 = time.perf_counter()
    env = host.Environment(path)
    sc = env.scene_arrays()
    env.close()
    t_cpu_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    g2 = rt.GpuScene(sc, precision=rt.FP32_FAST)
    torch.cuda.synchronize()
    t_arrays = time.perf_counter() - t0
    b = g2.render_region(W, H)
    g2.close()
    assert np.array_equal(a, b)
    assert (a.reshape(-1, 3).max(axis=1) > 0).mean() > 0.05
    out = {"obj_bytes": n_bytes, "faces": 991232, "host_read_and_scan_s": t_probe, "gpu_register_from_text_s": t_gpu,
           "gpu_ingest_ms": info["ms_ingest"], "bvh_ms": info["ms_build"], "host_loader_with_dedup_maps_s": t_cpu_load,
           "register_from_arrays_s": t_arrays, "test_writes_obj_s": t_write}
    print("OBJ_INGEST " + json.dumps(out))
    os.makedirs(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "r02_obj_ingest.json"), "w") as fh:
        json.dump(out, fh)
    assert t_gpu < t_cpu_load
