"""CPU: the C++ host mirror (host/state.cpp, host/gpu_worker.cpp) against the independent Python loader and
the reference's partition() semantics (master/main.go:54-91)."""
import json
import os

import numpy as np
import pytest

from conftest import EXAMPLE_SCENE, GOLDEN
from oracle import oracle as orc
from oracle import scene_loader as sl


@pytest.fixture(scope="module")
def host(built):
    from distributed_raytracer_b200 import host
    return host


def assert_same_scene(a, b):
    assert len(a["meshes"]) == len(b["meshes"])
    for ma, mb in zip(a["meshes"], b["meshes"]):
        for k in ("verts", "normals", "faces", "materials"):
            assert np.array_equal(np.asarray(ma[k]).reshape(-1), np.asarray(mb[k]).reshape(-1)), k
    assert np.array_equal(a["obj_pos"], b["obj_pos"]) and np.array_equal(a["obj_mesh"], b["obj_mesh"])
    assert np.array_equal(a["lights"], b["lights"])
    assert np.array_equal(a["camera"].view(np.uint64), b["camera"].view(np.uint64))


def test_environment_from_file_matches_python_loader(host):
    env = host.Environment(EXAMPLE_SCENE)
    sc = env.scene_arrays()
    assert_same_scene(sc, sl.environment_from_file(EXAMPLE_SCENE))
    assert sc["tan_half_fov"] == orc.go_tan(1.04719755 / 2.0)       # Go's math.Tan restated twice, independently


def test_loader_semantics_on_synthetic_obj(host, tmp_path):
    (tmp_path / "m.mtl").write_text("newmtl A\nKa 0.2 0.2 0.2\nKd 2.0 0.5 -1\nKs 0.1 0.2 0.3\nNs 4\n")
    (tmp_path / "m.obj").write_text(
        "mtllib m.mtl\no thing\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv -0.0 0 0\nv 0.1 0.2 0.3\nvn 0 0 2\nvn 0 3 0\n"
        "usemtl A\nf 1//1 2//1 3//2 4//2\ng other\nusemtl Missing\nf 5//2 2//1 6//1\nf -1//-1 -2//-2 -3//1\n")
    (tmp_path / "n.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    (tmp_path / "s.json").write_text(json.dumps({
        "objs": [{"model": "m.obj", "pos": {"x": 0, "y": 0, "z": 0}}, {"model": "n.obj", "pos": {"x": 1, "y": 2, "z": 3}},
                 {"model": "m.obj", "pos": {"x": -1, "y": 0, "z": 0}}],
        "lights": [{"pos": {"x": 0, "y": 0, "z": 5}, "col": {"r": 255, "g": 128, "b": 0}}],
        "cam": {"pos": {"x": 0, "y": 0, "z": 5}, "dir": {"x": 0.1, "y": 0.2, "z": -1}, "fov": 1.0}}))
    env = host.Environment(str(tmp_path / "s.json"))
    assert_same_scene(env.scene_arrays(), sl.environment_from_file(str(tmp_path / "s.json")))
    with pytest.raises(RuntimeError):
        host.Environment(str(tmp_path / "missing.json"))
    (tmp_path / "bad.json").write_text(json.dumps({"objs": [], "lights": [], "cam": {"pos": {"x": 0, "y": 0, "z": 0},
                                                                                    "dir": {"x": 0, "y": 1, "z": 0}, "fov": 1}}))
    with pytest.raises(RuntimeError, match="parallel"):
        host.Environment(str(tmp_path / "bad.json"))                 # camera.go:36-37


def test_camera_motion(host):
    env = host.Environment(EXAMPLE_SCENE)
    c0, _, _, _ = env.frame()
    env.cam_move(0.1, forward=True)                                   # camera.go:62-92
    c1, _, _, _ = env.frame()
    assert np.allclose(c1[0:3], c0[0:3] + 0.1 * c0[3:6], atol=1e-15)
    env.cam_yaw(0.3)
    env.cam_pitch(-0.2)
    c2, _, _, _ = env.frame()
    f, l, u = c2[3:6], c2[6:9], c2[9:12]
    assert np.allclose([f @ f, l @ l, u @ u], 1.0, atol=1e-12) and abs(f @ l) < 1e-12 and abs(l @ u) < 1e-12
    env.cam_yaw(0.0)                                                  # camera.go:131 theta % 2pi == 0 -> untouched
    c3, _, _, _ = env.frame()
    assert np.array_equal(c2, c3)


def rects_cover(rects, W, H):
    cover = np.zeros((W, H), dtype=np.int32)
    for (x, y, w, h) in rects:
        cover[x:x + w, y:y + h] += 1
    return (cover == 1).all()


@pytest.mark.parametrize("W,H,workers", [(320, 240, 1), (320, 240, 2), (320, 240, 4), (320, 240, 24), (321, 241, 7), (3840, 2160, 8),
                                          (60, 40, 64), (50, 50, 9)])
def test_partition_covers_the_screen_exactly_once(host, W, H, workers):
    rects = host.partition(W, H, workers)
    assert 1 <= len(rects) <= max(workers, 1)
    assert rects_cover(rects, W, H)


def test_partition_known_shapes(host):
    # master/main.go:54-91: bisect x first, then y, alternating; odd remainders go right / down
    assert host.partition(320, 240, 1) == [(0, 0, 320, 240)]
    assert host.partition(320, 240, 2) == [(0, 0, 160, 240), (160, 0, 160, 240)]
    assert host.partition(320, 240, 4) == [(0, 0, 160, 120), (0, 120, 160, 120), (160, 0, 160, 120), (160, 120, 160, 120)]
    assert host.partition(321, 241, 2) == [(0, 0, 160, 241), (160, 0, 161, 241)]
    assert host.partition(50, 50, 9) == [(0, 0, 50, 50)]              # <= widthKernel x heightKernel: not split (:66)
    assert host.partition(100, 40, 4) == [(0, 0, 50, 40), (50, 0, 50, 40)]   # height <= 50: only vertical cuts (:72)


def test_partition_balanced_tiles_the_screen_and_balances_cost():
    """partitionBalanced cuts by the cost of the previous frame.  The rectangles must tile the screen exactly, and
    on a frame whose hits sit in the middle (the BASELINE views) the most expensive rectangle must be far cheaper than with
    partition()'s geometric bisection (master/main.go:54-91)."""
    from distributed_raytracer_b200 import host
    W, H = 640, 360
    last = np.zeros((W, H, 3), dtype=np.uint8)                       # BulkTrace order: [i][j]
    ii, jj = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    blob = ((ii - 330) / 95.0) ** 2 + ((jj - 170) / 70.0) ** 2 < 1.0   # ~9 % of the pixels hit, off-centre
    last[blob] = (10, 200, 30)
    cost = 1.0 + 30.0 * blob

    def worst(rects):
        c = [cost[x:x + w, y:y + h].sum() for x, y, w, h in rects]
        return max(c) / (sum(c) / len(c))
    for workers in (2, 3, 5, 8, 16):
        rb = host.partition_balanced(W, H, workers, last, 30.0)
        rp = host.partition(W, H, workers)
        assert len(rb) == len(rp) == workers
        cover = np.zeros((W, H), dtype=np.int32)
        for x, y, w, h in rb:
            assert w > 0 and h > 0
            cover[x:x + w, y:y + h] += 1
        assert (cover == 1).all()
        assert worst(rb) <= 1.25, (workers, worst(rb))
        if workers >= 8:
            assert worst(rp) >= 2.0 * worst(rb), (workers, worst(rp), worst(rb))
    # without history it IS partition()
    black = np.zeros((W, H, 3), dtype=np.uint8)
    assert host.partition_balanced(W, H, 8, black, 30.0) == host.partition(W, H, 8)


def test_obj_text_host_half(tmp_path):
    """SURVEY s8 f3, host half of the OBJ ingest: bytes of the file, the material table (default material first, then first-use
    order, de-duplicated by value: mesh.go:151-163) and one usemtl entry per `usemtl` line.  (The parse itself runs on the GPU.)"""
    from distributed_raytracer_b200 import host
    n, n_mat, n_use = host.obj_text_probe(os.path.join(GOLDEN, "example", "suzanne.obj"))
    assert n == os.path.getsize(os.path.join(GOLDEN, "example", "suzanne.obj")) and n_mat == 2 and n_use == 1
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 1 0 0\nnewmtl b\nKd 1 0 0\nnewmtl c\nKd 0 1 0\n")
    (tmp_path / "m.obj").write_text("mtllib m.mtl\nv 0 0 0\nusemtl a\nusemtl b\n  usemtl c\nusemtl missing\n")
    n, n_mat, n_use = host.obj_text_probe(str(tmp_path / "m.obj"))
    assert n_use == 4 and n_mat == 3          # default (also what `missing` maps to), a == b by value, c
    with pytest.raises(RuntimeError):
        host.obj_text_probe(str(tmp_path / "nope.obj"))
