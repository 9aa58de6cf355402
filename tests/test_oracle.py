"""CPU: pins the oracle (oracle/rt_oracle.c) against the known answers of SURVEY.md Appendix C, checks that
its three candidate-enumeration modes agree, and covers the edge cases of the reference's geometry code."""
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, degenerate_scenes, random_soup_scene
from oracle import oracle as orc
from oracle import scene_loader as sl


def hx(x):
    return struct.pack(">d", float(x)).hex()


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "survey_appendix_c.json")) as fh:
        return json.load(fh)


def test_go_tan_known_answer(kat):
    assert hx(orc.go_tan(1.04719755 / 2.0)) == kat["tan_half_fov_hex"]
    assert orc.go_tan(0.0) == 0.0
    assert orc.go_tan(-0.3) == -orc.go_tan(0.3)
    for x in (0.1, 0.5, 0.785, 1.0, 1.3, 1.5):
        assert abs(orc.go_tan(x) - np.tan(x)) <= 4e-16 * abs(np.tan(x)) + 1e-300


def test_go_pow_integer_exponents_are_products():
    # math.Pow with an integer exponent is Frexp + square-and-multiply + Ldexp: compare with exact rational arithmetic
    from fractions import Fraction
    rng = np.random.default_rng(1)
    for x in list(rng.uniform(0.0, 1.0, 50)) + [0.5, 0.25, 1.0, 0.9999999999999999]:
        for n in (0, 1, 2, 3, 10, 17, 64):
            got = orc.go_pow(x, n)
            exact = Fraction(float(x)) ** n
            assert abs(Fraction(got) - exact) <= exact * Fraction(n + 1, 2 ** 52), (x, n)
    assert orc.go_pow(0.0, 10.0) == 0.0 and orc.go_pow(0.0, 0.0) == 1.0 and orc.go_pow(1.0, 1e9) == 1.0
    assert orc.go_pow(0.25, 0.5) == 0.5 and orc.go_pow(4.0, -0.5) == 0.5
    assert orc.go_pow(2.0, -2.0) == 0.25
    assert abs(orc.go_pow(0.7, 2.5) - 0.7 ** 2.5) < 1e-15


def test_appendix_c_known_answers(kat, example_scene, example_oracle):
    W, H = kat["width"], kat["height"]
    for px in kat["pixels"]:
        d = orc.primary_ray(px["i"], px["j"], W, H, example_scene["camera"])
        assert [hx(v) for v in d] == px["dir_hex"]
        for mode in (orc.BRUTE, orc.CULLED, orc.RTREE):
            if mode == orc.RTREE:
                example_oracle.build_rtrees()
            hit, col, face, obj, lit = example_oracle.trace_pixel(px["i"], px["j"], W, H, mode)
            assert face == px["face"] and hit == (px["face"] >= 0)
            if hit:
                assert [hx(c) for c in col] == px["colour_hex"]
                assert [(lit >> k) & 1 for k in range(3)] == px["lit"]
            assert [int(255 * c) for c in col] == px["rgb"]
            assert [int(65535 * c) >> 8 for c in col] == px["sdl"]
    d = orc.primary_ray(160, 120, W, H, example_scene["camera"])
    hit, ip, nrm, face, obj, mat = example_oracle.trace(example_scene["camera"][0:3], d, orc.BRUTE)
    assert hit and list(ip) == kat["hit_point_160_120"] and list(nrm) == kat["normal_160_120"]


def test_oracle_equals_independent_python_restatement(example_scene, example_oracle):
    """tests/pyref.py was written from the Go sources alone (scalar Python floats, brute force over the faces); the C oracle
    must agree with it bit for bit -- hit flag, face, lit mask, float64 colour, both quantisers -- on a few hundred pixels of
    C1: a seeded random sample, plus two scan lines and a column through the silhouette (where every branch of the
    intersection and shadow logic is exercised).  Two restatements by different routes that agree to the last bit leave
    little room for a transcription error in either; neither is the Go binary (parity stays unpinned)."""
    import pyref
    W, H = 320, 240
    S = pyref.Scene(example_scene)
    tan_half = orc.go_tan(float(example_scene["camera"][12]) / 2.0)
    rng = np.random.default_rng(20261018)
    pix = {(int(i), int(j)) for i, j in zip(rng.integers(0, W, 400), rng.integers(0, H, 400))}
    pix |= {(i, 120) for i in range(110, 210)} | {(i, 96) for i in range(120, 200, 2)} | {(160, j) for j in range(50, 190)}
    rgb255 = example_oracle.render_region(W, H, quant=orc.Q_RGB255)[0].reshape(W, H, 3)
    sdl = example_oracle.render_region(W, H, quant=orc.Q_RGBA16_SHR8)[0].reshape(W, H, 3)
    n_hit = 0
    for (i, j) in sorted(pix):
        hit, col, face, obj, lit = S.trace_pixel(i, j, W, H, tan_half)
        o_hit, o_col, o_face, o_obj, o_lit = example_oracle.trace_pixel(i, j, W, H, orc.BRUTE)
        assert (hit, face, lit) == (o_hit, o_face if o_hit else -1, o_lit if o_hit else 0), (i, j)
        if hit:
            n_hit += 1
            assert [hx(c) for c in col] == [hx(c) for c in o_col], (i, j)
        assert list(pyref.quantise_rgb(col)) == list(rgb255[i, j]), (i, j)
        assert list(pyref.quantise_rgba16_shr8(col)) == list(sdl[i, j]), (i, j)
    assert n_hit > 60 and len(pix) - n_hit > 60          # the sample really covers hits and misses


@pytest.mark.parametrize("kw", [dict(n_tris=150, seed=3, n_obj=3, n_lights=4), dict(n_tris=200, seed=4, with_normals=True, n_obj=2, n_lights=3)])
def test_oracle_equals_independent_python_restatement_on_soups(kw):
    """Same cross-check on random multi-object scenes: overlapping objects (the camera-distance metric of tracer.go:38 decides,
    also for shadow rays), flat normals (object.go:93-95) and interpolated ones, two materials with Ns = 10 and 3, up to 4
    lights -- every pixel of a 40x30 screen, all three candidate-enumeration modes of the oracle."""
    import pyref
    sc = random_soup_scene(**kw)
    W, H = 40, 30
    S = pyref.Scene(sc)
    O = orc.Scene(sc)
    O.build_rtrees()
    tan_half = orc.go_tan(float(sc["camera"][12]) / 2.0)
    n_hit = n_shadowed = 0
    for i in range(W):
        for j in range(H):
            hit, col, face, obj, lit = S.trace_pixel(i, j, W, H, tan_half)
            for mode in (orc.BRUTE, orc.CULLED, orc.RTREE):
                o_hit, o_col, o_face, o_obj, o_lit = O.trace_pixel(i, j, W, H, mode)
                assert hit == o_hit, (i, j, mode)
                if hit:
                    assert (face, obj, lit) == (o_face, o_obj, o_lit) and [hx(c) for c in col] == [hx(c) for c in o_col], (i, j, mode)
            n_hit += hit
            n_shadowed += hit and lit != (1 << len(sc["lights"])) - 1
    assert n_hit > 60 and n_shadowed > 10
    O.close()


@pytest.mark.parametrize("name", ["cube_edges_and_diagonals", "two_cubes_coplanar_fronts", "degenerate_triangles", "edge_on_triangle"])
def test_oracle_equals_independent_python_restatement_on_ties_and_degenerates(name):
    """Equal-distance ties (rays through shared edges, diagonals, corners: the first candidate in face order must win in both
    restatements), coplanar touching objects, zero-area triangles: every pixel, all three enumeration modes."""
    import pyref
    sc = degenerate_scenes()[name]
    S, O = pyref.Scene(sc), orc.Scene(sc)
    O.build_rtrees()
    tan_half = orc.go_tan(float(sc["camera"][12]) / 2.0)
    for (W, H) in [(16, 16), (17, 13)]:
        for i in range(W):
            for j in range(H):
                hit, col, face, obj, lit = S.trace_pixel(i, j, W, H, tan_half)
                for mode in (orc.BRUTE, orc.CULLED, orc.RTREE):
                    o_hit, o_col, o_face, o_obj, o_lit = O.trace_pixel(i, j, W, H, mode)
                    assert hit == o_hit, (i, j, mode)
                    if hit:
                        assert (face, obj, lit) == (o_face, o_obj, o_lit) and [hx(c) for c in col] == [hx(c) for c in o_col], (i, j, mode)
    O.close()


def test_hit_count_matches_survey(kat, example_oracle):
    _, _, face, st = example_oracle.render_region(320, 180, mode=orc.CULLED)
    assert st["hit_pixels"] == kat["stats_320x180"]["hit_pixels"] == int((face >= 0).sum())
    assert st["shadow_rays"] == 3 * st["hit_pixels"] and st["primary_rays"] == 320 * 180


@pytest.mark.parametrize("mode", ["CULLED", "BRUTE", "RTREE"])
def test_c1_golden_image_from_independent_restatement(example_oracle, mode):
    """The WHOLE C1 frame against tests/golden/c1_320x240_pyref.npz, which was produced by tests/golden/make_golden_pyref.py:
    its own scene/OBJ/MTL reader, its own math.Tan, the scalar Python arithmetic of tests/pyref.py -- nothing of the oracle
    or of the product.  Face map, float64 colour bits, lit masks and both quantisations must agree on every pixel."""
    g = np.load(os.path.join(GOLDEN, "c1_320x240_pyref.npz"))
    rgb, col, face, st = example_oracle.render_region(320, 240, mode=getattr(orc, mode))
    assert np.array_equal(face, g["face"].astype(np.int32))
    hit = g["hit_index"]
    assert st["hit_pixels"] == len(hit) == 5820
    assert np.array_equal(col[hit].view(np.uint64), g["colour_bits"])
    assert np.array_equal(rgb[hit], g["rgb"]) and not rgb[face < 0].any()
    sdl, _, _, _ = example_oracle.render_region(320, 240, mode=getattr(orc, mode), quant=orc.Q_RGBA16_SHR8, want_colour=False)
    assert np.array_equal(sdl[hit], g["sdl"])
    assert (sdl != rgb).any()          # the two quantisers differ


def test_c2_golden_sample_from_independent_restatement(example_scene):
    """C2 (1920x1080, lights[0] only): every 97th pixel of the BulkTrace order against the independent restatement."""
    g = np.load(os.path.join(GOLDEN, "c2_1080p_sample_pyref.npz"))
    sc = dict(example_scene)
    sc["lights"] = example_scene["lights"][:1]
    S = orc.Scene(sc)
    rgb, col, face, st = S.render_region(1920, 1080, mode=orc.CULLED, n_threads=0)
    S.close()
    idx = g["index"]
    assert np.array_equal(face[idx], g["face"].astype(np.int32))
    assert np.array_equal(col[idx].view(np.uint64), g["colour_bits"])
    assert np.array_equal(rgb[idx], g["rgb"])
    assert (g["face"] >= 0).sum() > 2000


def test_c3_golden_sample_from_independent_restatement(example_scene):
    """Config 3 (991 232 faces, 3840x2160, lights[0..1]): 404 pixels rendered by tests/golden/make_golden_pyref_c3.py -- its own
    subdivision of Suzanne, its own reader and math.Tan, tests/pyref.py arithmetic over ALL faces -- against the oracle on the
    mesh workloads.subdivide makes: the face index (so the two subdivisions number and place the faces alike), the float64
    colour bits, the lit masks and both quantisations."""
    from distributed_raytracer_b200 import workloads
    g = np.load(os.path.join(GOLDEN, "c3_4k_sample_pyref.npz"))
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 5)]
    sc["lights"] = example_scene["lights"][:2]
    S = orc.Scene(sc)
    assert (g["face"] >= 0).sum() >= 200
    for k, (i, j) in enumerate(zip(g["pixel_i"], g["pixel_j"])):
        hit, col, face, obj, lit = S.trace_pixel(int(i), int(j), 3840, 2160, mode=orc.CULLED)
        assert face == int(g["face"][k]), (i, j)
        if hit:
            assert np.array_equal(col.view(np.uint64), g["colour_bits"][k]), (i, j)
            assert lit == int(g["lit"][k])
            assert tuple(int(255 * c) for c in col) == tuple(int(v) for v in g["rgb"][k])
    S.close()


def test_golden_generator_loader_agrees_with_the_oracle_loader(example_scene):
    """The golden generator reads scene.json / suzanne.obj / suzanne.mtl with its own code; the per-face geometry it feeds
    its renderer must be the geometry the oracle's loader produces (different indexing: it does not deduplicate)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_pyref", os.path.join(GOLDEN, "make_golden_pyref.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    sc = mg.read_obj_scene(os.path.join(GOLDEN, "example", "scene.json"))
    m, mo = sc["meshes"][0], example_scene["meshes"][0]
    v, n, f = np.array(m["verts"]), np.array(m["normals"]), np.array(m["faces"])
    vo, no, fo = np.asarray(mo["verts"]).reshape(-1, 3), np.asarray(mo["normals"]).reshape(-1, 3), np.asarray(mo["faces"]).reshape(-1, 7)
    assert f.shape == fo.shape == (968, 7)
    assert np.array_equal(v[f[:, 0:3]].view(np.uint64), vo[fo[:, 0:3]].view(np.uint64))
    assert np.array_equal(n[f[:, 3:6]].view(np.uint64), no[fo[:, 3:6]].view(np.uint64))
    assert np.array_equal(np.array(m["materials"])[f[:, 6]].view(np.uint64), np.asarray(mo["materials"]).reshape(-1, 10)[fo[:, 6]].view(np.uint64))
    assert np.array_equal(np.array(sc["camera"]).view(np.uint64), np.asarray(example_scene["camera"], dtype=np.float64).view(np.uint64))
    assert np.array_equal(np.array(sc["lights"]).view(np.uint64), np.asarray(example_scene["lights"], dtype=np.float64).reshape(-1, 6).view(np.uint64))
    assert mg.go_tan(sc["camera"][12] / 2.0) == orc.go_tan(sc["camera"][12] / 2.0)


@pytest.mark.parametrize("kw", [dict(n_tris=60, seed=3), dict(n_tris=200, seed=4, with_normals=True),
                                dict(n_tris=120, seed=5, n_obj=3, n_lights=3), dict(n_tris=1, seed=6), dict(n_tris=2, seed=7)])
def test_modes_agree_on_random_scenes(kw):
    sc = random_soup_scene(**kw)
    S = orc.Scene(sc)
    ref = S.render_region(64, 48, mode=orc.BRUTE)
    for mode in (orc.CULLED, orc.RTREE):
        got = S.render_region(64, 48, mode=mode)
        assert np.array_equal(got[2], ref[2])
        assert np.array_equal(got[1].view(np.uint64), ref[1].view(np.uint64))
        assert np.array_equal(got[0], ref[0])
    assert (ref[2] >= 0).any()


def test_region_layout_and_sampling(example_oracle):
    W, H = 96, 72
    full_rgb, _, full_face, _ = example_oracle.render_region(W, H, layout=orc.L_COLMAJOR)
    x, y, w, h = 30, 20, 41, 33
    rgb, _, face, st = example_oracle.render_region(W, H, x, y, w, h, layout=orc.L_COLMAJOR)
    full = full_face.reshape(W, H)                      # column-major: index i*H + j  (distributed/main.go:82)
    assert np.array_equal(face.reshape(w, h), full[x:x + w, y:y + h])
    rgb_r, _, face_r, _ = example_oracle.render_region(W, H, x, y, w, h, layout=orc.L_ROWMAJOR)
    assert np.array_equal(face_r.reshape(h, w), face.reshape(w, h).T)
    assert np.array_equal(rgb_r.reshape(h, w, 3), rgb.reshape(w, h, 3).transpose(1, 0, 2))
    # bounded sample: only every 4th region-linear pixel is traced
    _, _, f4, st4 = example_oracle.render_region(W, H, x, y, w, h, step=4, phase=1)
    idx = np.arange(w * h)
    assert np.array_equal(f4[idx % 4 == 1], face[idx % 4 == 1]) and (f4[idx % 4 != 1] == -1).all()
    assert st4["primary_rays"] == (idx % 4 == 1).sum()
    # single- and multi-threaded renders agree
    a = example_oracle.render_region(W, H, n_threads=1)
    b = example_oracle.render_region(W, H, n_threads=4)
    assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


def test_triangle_intersection_edge_cases():
    tri = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    hit, ip, bary, t = orc.triangle_intersection(tri, [0.25, 0.25, 1.0], [0, 0, -1.0])
    assert hit and t == 1.0 and list(ip) == [0.25, 0.25, 0.0] and list(bary) == [0.5, 0.25, 0.25]
    # two-sided (triangle.go:46 only needs incidence != 0)
    assert orc.triangle_intersection(tri, [0.25, 0.25, -1.0], [0, 0, 1.0])[0]
    # inclusive edges and vertices (triangle.go:53,58,63)
    assert orc.triangle_intersection(tri, [0.5, 0.5, 1.0], [0, 0, -1.0])[0]
    assert orc.triangle_intersection(tri, [0.0, 0.0, 1.0], [0, 0, -1.0])[0]
    assert orc.triangle_intersection(tri, [1.0, 0.0, 1.0], [0, 0, -1.0])[0]
    assert not orc.triangle_intersection(tri, [0.5000001, 0.5000001, 1.0], [0, 0, -1.0])[0]
    # behind the origin (triangle.go:68), on the origin (t == 0 is accepted), parallel (incidence == 0)
    assert not orc.triangle_intersection(tri, [0.25, 0.25, -1.0], [0, 0, -1.0])[0]
    assert orc.triangle_intersection(tri, [0.25, 0.25, 0.0], [0, 0, -1.0])[0]
    assert not orc.triangle_intersection(tri, [0.25, 0.25, 1.0], [1.0, 0, 0])[0]
    # degenerate triangle: incidence == 0 -> never hit
    assert not orc.triangle_intersection([0, 0, 0, 1, 0, 0, 2, 0, 0], [0.5, 0, 1.0], [0, 0, -1.0])[0]


def test_box_intersect_reference_predicate():
    lo, hi = [0, 0, 0], [1, 1, 1]
    assert orc.box_intersect(lo, hi, [0.5, 0.5, 5], [0, 0, -1])
    assert not orc.box_intersect(lo, hi, [0.5, 0.5, 5], [0, 0, 1])      # behind: t >= 0 required (box.go:42)
    assert orc.box_intersect(lo, hi, [0.5, 0.5, 0.5], [0, 0, 1])         # origin inside
    assert orc.box_intersect(lo, hi, [1.0, 0.5, 5], [0, 0, -1])          # inclusive bounds (box.go:47)
    assert not orc.box_intersect(lo, hi, [1.0000001, 0.5, 5], [0, 0, -1])
    assert orc.box_intersect(lo, hi, [-1, -1, -1], [1, 1, 1])


def test_loader_example_scene(example_scene):
    m = example_scene["meshes"][0]
    assert m["faces"].shape == (968, 7) and m["verts"].shape == (505, 3) and m["normals"].shape == (507, 3)
    # float32 parse then widen (mesh.go:176), MTL floats likewise (colour.go:33-35)
    assert np.array_equal(m["verts"], m["verts"].astype(np.float32).astype(np.float64))
    assert list(m["materials"][0]) == [float(np.float32(0.1))] * 3 + [float(np.float32(0.8))] * 3 + [0.5] * 3 + [10.0]
    # stored normals are normalised (mesh.go:202)
    assert np.allclose(np.sqrt((m["normals"] ** 2).sum(axis=1)), 1.0, atol=1e-15)
    assert np.array_equal(example_scene["lights"][:, 3:], np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1.0]]))
    cam = example_scene["camera"]
    assert list(cam[3:12]) == [0, 0, -1, 1, -0.0, 0, 0, 1, 0] and np.signbit(cam[7])   # left = (1,-0,0), Appendix B-8


def test_loader_quads_groups_defaults(tmp_path):
    (tmp_path / "m.mtl").write_text("newmtl A\nKa 0.2 0.2 0.2\nKd 2.0 0.5 -1\nKs 0 0 0\nNs 4\n")
    (tmp_path / "m.obj").write_text(
        "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv -0.0 0 0\nv 0.1 0.2 0.3\n"
        "usemtl A\nf 1 2 3 4\nusemtl Missing\nf 5 2 6\nf -1 -2 -3\n")
    (tmp_path / "s.json").write_text(json.dumps({
        "objs": [{"model": "m.obj", "pos": {"x": 0, "y": 0, "z": 0}}, {"model": "m.obj", "pos": {"x": 1, "y": 2, "z": 3}}],
        "lights": [{"pos": {"x": 0, "y": 0, "z": 5}, "col": {"r": 255, "g": 128, "b": 0}}],
        "cam": {"pos": {"x": 0, "y": 0, "z": 5}, "dir": {"x": 0, "y": 0, "z": -1}, "fov": 1.0}}))
    sc = sl.environment_from_file(str(tmp_path / "s.json"))
    assert len(sc["meshes"]) == 1 and list(sc["obj_mesh"]) == [0, 0]          # meshes are shared per model path
    m = sc["meshes"][0]
    # quad -> (v0 v1 v2) + (v2 v3 v0); -0.0 dedups onto +0.0 (Go map keys); negative indices are relative
    assert m["faces"][:, 0:3].tolist() == [[0, 1, 2], [2, 3, 0], [0, 1, 4], [4, 0, 3]]
    assert m["normals"].shape[0] == 0
    assert m["faces"][:, 6].tolist() == [0, 0, 1, 1]
    assert list(m["materials"][0]) == [float(np.float32(0.2))] * 3 + [1.0, 0.5, 0.0] + [0.0] * 3 + [4.0]      # clamped to [0,1]
    assert list(m["materials"][1]) == [16 / 255.0] * 3 + [1.0] * 3 + [0.0] * 3 + [0.0]                         # mesh.go:151
    assert list(sc["lights"][0]) == [0, 0, 5, 1.0, 128 / 255.0, 0.0]
    with pytest.raises(ValueError):
        sl.new_camera([0, 0, 0], [0, 1, 0], 1.0)                                # camera.go:36-37
