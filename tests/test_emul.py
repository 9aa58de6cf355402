"""CPU: replays the device code (the RT_HD functions the sm_100a kernels are made of: LBVH builder bodies,
fp64-exact and fp32-fast traversal, fused shadow+shade, quantisers, ray<->pixel maps) through tests/emul and
compares it with the oracle.  This is what can be proven about the kernels without a GPU; the -m gpu tests
repeat the same comparisons through the C ABI on the real device."""
import numpy as np
import pytest

import emul_build as eb
from conftest import box_predicate_scene, degenerate_scenes, random_soup_scene
from oracle import oracle as orc

import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import workloads


def oracle_render(sc, W, H, **kw):
    S = orc.Scene(sc)
    return S.render_region(W, H, mode=orc.CULLED, **kw)


def check_exact(sc, W, H, max_leaf=4, flags=0):
    rgb_o, col_o, face_o, st = oracle_render(sc, W, H)
    E = eb.EmulScene(sc, rt.FP64_EXACT, max_leaf)
    rc = E.bvh_check()
    assert rc[0] == 0, "BVH structure invalid: code %d" % rc[0]
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H, flags=flags)
    assert np.array_equal(face, face_o)
    assert np.array_equal(col.view(np.uint64), col_o.view(np.uint64))      # bit-exact float64 colours
    assert np.array_equal(rgb, rgb_o)
    assert int(cnt[2]) == st["hit_pixels"]
    E.close()
    return st


def check_fast(sc, W, H, max_leaf=4):
    rgb_o, col_o, face_o, st = oracle_render(sc, W, H)
    E = eb.EmulScene(sc, rt.FP32_FAST, max_leaf)
    assert E.bvh_check()[0] == 0
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H)
    E.close()
    within = (np.abs(rgb.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean()
    same_face = (face == face_o).mean()
    return within, same_face


@pytest.mark.parametrize("max_leaf", [1, 4, 8])
def test_exact_replay_bit_identical_on_example(example_scene, max_leaf):
    st = check_exact(example_scene, 320, 240, max_leaf)
    assert st["hit_pixels"] == 5820


def test_exact_replay_brute_force_flag(example_scene):
    check_exact(example_scene, 96, 72, flags=rt.FLAG_BRUTE_FORCE)


@pytest.mark.parametrize("kw", [dict(n_tris=300, seed=11), dict(n_tris=500, seed=12, with_normals=True),
                                dict(n_tris=150, seed=13, n_obj=3, n_lights=4), dict(n_tris=1, seed=14, half=1.2, extent=0.2), dict(n_tris=2, seed=15, half=1.2, extent=0.2),
                                dict(n_tris=3, seed=16, half=1.2, extent=0.2), dict(n_tris=40, seed=17, half=1.5)])
def test_exact_replay_bit_identical_on_random_scenes(kw):
    sc = random_soup_scene(**kw)
    st = check_exact(sc, 80, 60)
    assert st["hit_pixels"] > 0


def test_exact_replay_on_subdivided_suzanne(example_scene):
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 2)]
    assert sc["meshes"][0]["faces"].shape[0] == 968 * 16
    check_exact(sc, 160, 120)


def test_exact_replay_with_duplicate_triangles():
    # identical triangles => identical Morton keys (index tie-break in the Karras tree) and exact-distance
    # ties that must resolve to the lowest face index
    sc = random_soup_scene(n_tris=20, seed=21)
    m = sc["meshes"][0]
    m["verts"] = np.concatenate([m["verts"]] * 4, axis=0)
    f = m["faces"].copy()
    m["faces"] = np.concatenate([f, f + [60, 60, 60, 0, 0, 0, 0], f + [120, 120, 120, 0, 0, 0, 0], f + [180, 180, 180, 0, 0, 0, 0]]).astype(np.uint32)
    rgb_o, col_o, face_o, st = oracle_render(sc, 64, 48)
    assert st["nearest_ties"] > 0 and face_o.max() < 20
    check_exact(sc, 64, 48)
    check_exact(sc, 64, 48, max_leaf=1)


def soup_config_scene(n_tris, example_scene, n_lights=8, angle_deg=0.0):
    """Config-4-shaped scene at test size: random soup (no normals, default material), ring of white lights, orbit camera."""
    sc = dict(example_scene)
    sc["meshes"] = [workloads.triangle_soup(n_tris, half=0.05)]
    sc["lights"] = workloads.ring_lights(n_lights)
    sc["camera"] = workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(angle_deg), 0.5, 1.04719755)
    return sc


def test_exact_replay_soup_flat_normals_eight_lights(example_scene):
    sc = soup_config_scene(5000, example_scene, angle_deg=33.0)
    st = check_exact(sc, 120, 68)
    assert st["hit_pixels"] > 500 and st["shadow_rays"] == 8 * st["hit_pixels"]


@pytest.mark.parametrize("name", ["cube_edges_and_diagonals", "two_cubes_coplanar_fronts"])
def test_exact_replay_exact_edge_and_corner_hits(name):
    """Rays through shared edges, face diagonals and corners of axis-aligned cubes (equal-distance ties, zero-thickness BVH
    boxes, two objects whose front faces are coplanar and touch): the replayed kernels pick the same face as the oracle."""
    sc = degenerate_scenes()[name]
    for (W, H) in [(16, 16), (32, 32), (17, 13)]:
        st = check_exact(sc, W, H)
        assert st["hit_pixels"] > 0
    S = orc.Scene(sc)
    assert S.render_region(16, 16, mode=orc.BRUTE)[3]["nearest_ties"] > 0      # the scene really produces exact ties
    S.close()


@pytest.mark.parametrize("name", ["no_objects", "mesh_without_faces", "single_triangle", "degenerate_triangles", "edge_on_triangle"])
@pytest.mark.parametrize("W,H", [(1, 1), (3, 5), (64, 48)])
def test_exact_replay_degenerate_inputs(name, W, H):
    sc = degenerate_scenes()[name]
    st = check_exact(sc, W, H)
    if name in ("no_objects", "mesh_without_faces"):
        assert st["hit_pixels"] == 0
    if name == "single_triangle" and W == 64:
        assert st["hit_pixels"] > 50
    within, same_face = check_fast(sc, W, H)          # fp32 traversal may decide a silhouette pixel differently
    assert within >= 0.99 and same_face >= 0.99


def test_orbit_camera_matches_reference_basis(example_scene):
    # workloads.orbit_camera builds the basis exactly like state.NewCamera (camera.go:35-44)
    from oracle import scene_loader as sl
    cam = workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(57.0), 0.3, 1.0)
    ref = sl.new_camera(cam[0:3], np.array([1.0, 1.0, -1.0]) - cam[0:3], 1.0)
    assert np.array_equal(cam.view(np.uint64), ref.view(np.uint64))
    assert np.array_equal(workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, 0.0, 0.0, 1.04719755)[0:3], example_scene["camera"][0:3])


def test_fast_replay_within_tolerance(example_scene):
    within, same_face = check_fast(example_scene, 320, 240)
    assert within >= 0.999 and same_face >= 0.999       # north_star: <= 1 LSB on >= 99.9 % of the pixels


@pytest.mark.parametrize("kw", [dict(n_tris=400, seed=31, with_normals=True), dict(n_tris=150, seed=32, n_obj=3, n_lights=4)])
def test_fast_replay_random_scenes(kw):
    within, same_face = check_fast(random_soup_scene(**kw), 96, 72)
    assert within >= 0.995 and same_face >= 0.995       # tiny random soups have many silhouette pixels


def test_fast_replay_with_fp32_shading_option(example_scene, monkeypatch):
    """RT_B200_SHADE32=1 (opt-in): FP32_FAST with the shading in fp32 too still meets the tolerance of north_star (<= 1 LSB
    on >= 99.9 % of the pixels); what it gives up is the bit-identical colours of the default fp64 shading."""
    monkeypatch.setenv("RT_B200_SHADE32", "1")
    within, same_face = check_fast(example_scene, 320, 240)
    assert within >= 0.999 and same_face >= 0.999
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 2)]
    within, same_face = check_fast(sc, 480, 270)
    assert within >= 0.999 and same_face >= 0.999
    for kw in (dict(n_tris=400, seed=31, with_normals=True), dict(n_tris=150, seed=32, n_obj=3, n_lights=4)):
        within, same_face = check_fast(random_soup_scene(**kw), 160, 120)
        assert within >= 0.99 and same_face >= 0.99
    monkeypatch.setenv("RT_B200_SHADE32", "0")
    assert check_fast(example_scene, 320, 240)[0] == 1.0          # the default stays bit-identical on C1


def test_region_layout_quantiser(example_scene):
    W, H, x, y, w, h = 96, 72, 30, 20, 41, 33
    S = orc.Scene(example_scene)
    E = eb.EmulScene(example_scene, rt.FP64_EXACT)
    for flags, layout, quant in [(0, orc.L_COLMAJOR, orc.Q_RGB255), (rt.LAYOUT_ROWMAJOR, orc.L_ROWMAJOR, orc.Q_RGB255),
                                 (rt.QUANT_RGBA16_SHR8, orc.L_COLMAJOR, orc.Q_RGBA16_SHR8)]:
        rgb_o, col_o, face_o, _ = S.render_region(W, H, x, y, w, h, layout=layout, quant=quant)
        rgb, col, face, obj, lit, cnt = E.render(example_scene["camera"], example_scene["lights"], W, H, x, y, w, h, flags=flags)
        assert np.array_equal(rgb, rgb_o) and np.array_equal(face, face_o)
    # odd screen size: integer half sizes (tracer.go:16), ragged 8x4 blocks
    rgb_o, _, face_o, _ = S.render_region(101, 67)
    rgb, _, face, _, _, _ = E.render(example_scene["camera"], example_scene["lights"], 101, 67)
    assert np.array_equal(rgb, rgb_o) and np.array_equal(face, face_o)
    E.close()


def numpy_untile(all_tiles, W, H, tw, th, n_ranks):
    """Mirror of k_untile (render.cu), row-major output."""
    tiles_x = (W + tw - 1) // tw
    per_rank = rt.tiles_per_rank(W, H, tw, th, n_ranks)
    tiles = all_tiles.reshape(n_ranks, per_rank, th, tw, 3)
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ty = jj // th
    t = ty * tiles_x + (ii // tw - 3 * ty) % tiles_x          # csrc/rt_types.h: tile_id()
    return tiles[t % n_ranks, t // n_ranks, jj % th, ii % tw]


@pytest.mark.parametrize("W,H,tw,th,n", [(96, 72, 32, 16, 3), (100, 70, 16, 8, 2), (64, 48, 8, 4, 8)])
def test_tile_sharding_maps(example_scene, W, H, tw, th, n):
    E = eb.EmulScene(example_scene, rt.FP64_EXACT)
    ref, _, _, _, _, _ = E.render(example_scene["camera"], example_scene["lights"], W, H, flags=rt.LAYOUT_ROWMAJOR)
    per = rt.tiles_per_rank(W, H, tw, th, n)
    bufs = []
    for r in range(n):
        out = np.zeros(per * tw * th * 3, dtype=np.uint8)
        E.render_tiles(example_scene["camera"], example_scene["lights"], W, H, tw, th, r, n, out)
        bufs.append(out)
    img = numpy_untile(np.concatenate(bufs), W, H, tw, th, n)
    assert np.array_equal(img, ref.reshape(H, W, 3))
    E.close()


def test_bvh_builder_structure():
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 5, 17, 64, 257, 1500]:
        sc = random_soup_scene(n_tris=n, seed=100 + n)
        for max_leaf in (1, 4, 16):
            E = eb.EmulScene(sc, rt.FP32_FAST, max_leaf)
            rc, n_nodes, depth, n_leaves = E.bvh_check()
            assert rc == 0, (n, max_leaf, rc)
            assert n_nodes <= max(n - 1, 0) and depth <= 94
            if max_leaf == 1 and n > 1:
                # a 4-wide node has 2..4 children: (n-1)/3 <= nodes <= n-1; every face is its own leaf
                assert (n - 1 + 2) // 3 <= n_nodes <= n - 1 and n_leaves == n
            E.close()
    # Morton code: bit interleave of 21-bit cells, x most significant
    L = eb.lib()
    assert L.emul_morton63(0.0, 0.0, 0.0) == 0
    assert L.emul_morton63(1.0, 1.0, 1.0) == (1 << 63) - 1
    assert L.emul_morton63(0.5, 0.0, 0.0) == 1 << 62 and L.emul_morton63(0.0, 0.5, 0.0) == 1 << 61 and L.emul_morton63(0.0, 0.0, 0.5) == 1 << 60
    a = rng.uniform(0, 1, (200, 3))
    for p in a:
        q = (np.minimum(p * 2 ** 21, 2 ** 21 - 1)).astype(np.int64)
        want = 0
        for b in range(21):
            want |= ((int(q[0]) >> b) & 1) << (3 * b + 2) | ((int(q[1]) >> b) & 1) << (3 * b + 1) | ((int(q[2]) >> b) & 1) << (3 * b)
        assert L.emul_morton63(*p) == want


def test_go_math_device_functions_match_oracle():
    """go_max / go_min / go_pow of csrc/rt_exact.cuh (device code) against the oracle's restatement of Go's
    math.Max / math.Min / math.Pow, bit for bit, including every special case (NaN, +-Inf, signed zeros)."""
    import struct
    L, O = eb.lib(), orc.lib()
    bits = lambda v: struct.pack(">d", v)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 1e-300, -1e-300, 1.7976931348623157e308, -1.7976931348623157e308,
            float("inf"), float("-inf"), float("nan"), 0.3, 10.0, 3.0, -2.0, 2.5, 1e308]
    for x in vals:
        for y in vals:
            assert bits(L.emul_go_max(x, y)) == bits(O.orc_go_max(x, y)) or (np.isnan(L.emul_go_max(x, y)) and np.isnan(O.orc_go_max(x, y))), (x, y)
            assert bits(L.emul_go_min(x, y)) == bits(O.orc_go_min(x, y)) or (np.isnan(L.emul_go_min(x, y)) and np.isnan(O.orc_go_min(x, y))), (x, y)
            a, b = L.emul_go_pow(x, y), O.orc_go_pow(x, y)
            assert bits(a) == bits(b) or (np.isnan(a) and np.isnan(b)), (x, y, a, b)
    rng = np.random.default_rng(9)
    for x in rng.uniform(0, 1.0000001, 300):
        for n in (0.0, 1.0, 2.0, 10.0, 33.0, 0.5, 200.0):
            assert bits(L.emul_go_pow(x, n)) == bits(O.orc_go_pow(x, n))


# ---- round 2: primary-ray culling rectangle (csrc/rt_cull.h) and the separate any-hit hierarchy ----------------------
@pytest.mark.parametrize("precision", [rt.FP64_EXACT, rt.FP32_FAST])
def test_cull_rectangle_never_changes_a_pixel(example_scene, precision):
    """The projected-root-box rectangle may only skip rays that cannot hit: frames with and without it are identical, for
    the example view, for cameras orbiting the object (incl. views where the object is cut by the screen edge or lies
    behind the camera), for sub-regions and for tile mode."""
    E = eb.EmulScene(example_scene, precision)
    lights = example_scene["lights"]
    views = [(example_scene["camera"], 160, 120)]
    for k, (radius, elev) in enumerate([(6.0, 0.3), (2.5, -0.4), (1.2, 0.9), (0.4, 0.0), (9.0, 1.2)]):
        views.append((workloads.orbit_camera([1.0, 1.0, -1.0], radius, np.deg2rad(40.0 * k + 10.0), elev, 1.0), 96 + 8 * k, 72 + 5 * k))
    # looking away from the object: every corner behind the camera -> no culling information, still an all-black frame
    away = np.array(example_scene["camera"], dtype=np.float64).copy()
    away[3:6] = -away[3:6]
    from oracle import scene_loader as sl
    views.append((sl.new_camera(away[0:3], away[3:6], float(away[12])), 64, 48))
    for cam, W, H in views:
        a = E.render(cam, lights, W, H, flags=rt.FLAG_NO_CULL)
        b = E.render(cam, lights, W, H)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
        assert int(a[5][2]) == int(b[5][2])                                     # same number of hit pixels
        x, y, w, h = W // 3, H // 4, W // 2, H // 2
        a = E.render(cam, lights, W, H, x, y, w, h, flags=rt.FLAG_NO_CULL)
        b = E.render(cam, lights, W, H, x, y, w, h)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    cam = example_scene["camera"]
    per = rt.tiles_per_rank(96, 72, 16, 8, 3)
    for r in range(3):
        t0, t1 = np.zeros(per * 16 * 8 * 3, dtype=np.uint8), np.zeros(per * 16 * 8 * 3, dtype=np.uint8)
        E.render_tiles(cam, lights, 96, 72, 16, 8, r, 3, t0, flags=rt.FLAG_NO_CULL)
        E.render_tiles(cam, lights, 96, 72, 16, 8, r, 3, t1)
        assert np.array_equal(t0, t1)
    E.close()


def test_cull_rectangle_skips_most_of_the_example_frame(example_scene):
    """C1/C3 view: ~10 % of the pixels hit; the rectangle must cut the primary-ray work accordingly (node visits counted by
    the replay: rays outside the rectangle are never generated)."""
    E = eb.EmulScene(example_scene, rt.FP32_FAST)
    a = E.render(example_scene["camera"], example_scene["lights"], 320, 240, flags=rt.FLAG_NO_CULL)
    b = E.render(example_scene["camera"], example_scene["lights"], 320, 240)
    assert np.array_equal(a[0], b[0])
    assert int(b[5][0]) <= int(a[5][0])                 # never more node visits
    E.close()


def test_cull_rectangle_multi_object_and_degenerate():
    for kw in (dict(n_tris=150, seed=13, n_obj=3, n_lights=4), dict(n_tris=1, seed=14, half=1.2, extent=0.2)):
        sc = random_soup_scene(**kw)
        E = eb.EmulScene(sc, rt.FP64_EXACT)
        a = E.render(sc["camera"], sc["lights"], 96, 72, flags=rt.FLAG_NO_CULL)
        b = E.render(sc["camera"], sc["lights"], 96, 72)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        E.close()
    for name, sc in degenerate_scenes().items():
        E = eb.EmulScene(sc, rt.FP64_EXACT)
        a = E.render(sc["camera"], sc["lights"], 64, 48, flags=rt.FLAG_NO_CULL)
        b = E.render(sc["camera"], sc["lights"], 64, 48)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]), name
        E.close()


@pytest.mark.parametrize("leaf_any", [1, 2, 8])
def test_any_hit_hierarchy_with_its_own_leaf_size(example_scene, leaf_any):
    """FP32_FAST shadow rays walk a second hierarchy over the same triangles (RT_B200_MAX_LEAF_SHADE): the lit masks and
    colours must not depend on its leaf size."""
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 1)]
    E0 = eb.EmulScene(sc, rt.FP32_FAST, 4, 4)
    E1 = eb.EmulScene(sc, rt.FP32_FAST, 4, leaf_any)
    assert E1.bvh_check()[0] == 0
    a = E0.render(sc["camera"], sc["lights"], 240, 180)
    b = E1.render(sc["camera"], sc["lights"], 240, 180)
    assert np.array_equal(a[2], b[2])                                            # primary hits: same hierarchy
    assert (a[4] == b[4]).mean() >= 0.9999                                       # lit masks
    assert (a[0] == b[0]).all(axis=1).mean() >= 0.9999
    rgb_o, _, face_o, _ = oracle_render(sc, 240, 180)
    assert (np.abs(b[0].astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean() >= 0.999
    E0.close(); E1.close()


@pytest.mark.parametrize("precision", [rt.FP64_EXACT, rt.FP32_FAST])
def test_object_with_nan_position_is_absent(precision):
    """An order that does not list a registered object (environment.go:73-98 rebuilds the object tree from the diff alone) hands
    the backend a NaN position for it: the frame must be the frame of the scene WITHOUT that object."""
    sc = random_soup_scene(n_tris=150, seed=13, n_obj=3, n_lights=4)
    without = dict(sc, obj_pos=np.asarray(sc["obj_pos"])[[0, 2]], obj_mesh=np.asarray(sc["obj_mesh"])[[0, 2]])
    rgb_o, col_o, face_o, _ = oracle_render(without, 96, 72)
    pos = np.asarray(sc["obj_pos"], dtype=np.float64).copy()
    pos[1] = np.nan
    E = eb.EmulScene(dict(sc, obj_pos=pos), precision)
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], 96, 72)
    E.close()
    assert not (obj == 1).any() and (obj == 2).any()
    if precision == rt.FP64_EXACT:
        assert np.array_equal(face, face_o) and np.array_equal(col.view(np.uint64), col_o.view(np.uint64)) and np.array_equal(rgb, rgb_o)
    else:
        assert (face == face_o).mean() >= 0.995 and (np.abs(rgb.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean() >= 0.995


def test_exact_replay_applies_the_reference_object_level_predicate():
    """tracer.go:32 + box.go:29-69: the reference searches an object only if its NON-conservative box predicate accepts the
    object's bounds.  In this stress scene the predicate rejects the bounds for one shadow ray that really crosses them; the
    replay of the FP64_EXACT kernels follows the reference-shaped walk of the oracle bit for bit, while a conservative search
    (the oracle's distance-pruning mode) finds the occluder the reference never looks at."""
    sc, W, H = box_predicate_scene()
    S = orc.Scene(sc)
    S.build_rtrees()
    rgb_r, col_r, face_r, _ = S.render_region(W, H, mode=orc.RTREE)
    rgb_c, col_c, face_c, _ = S.render_region(W, H, mode=orc.CULLED)
    E = eb.EmulScene(sc, rt.FP64_EXACT, 4)
    rgb, col, face, obj, lit, cnt = E.render(sc["camera"], sc["lights"], W, H)
    E.close(); S.close()
    assert np.array_equal(face, face_r) and np.array_equal(col.view(np.uint64), col_r.view(np.uint64)) and np.array_equal(rgb, rgb_r)
    assert not np.array_equal(col_c.view(np.uint64), col_r.view(np.uint64))      # the fixture does exercise the corner


def test_quantised_node_experiment_renders_the_identical_frame(example_scene):
    """-DRT_NODEQ=1 (64-byte nodes on a 16-bit grid, measured slower on B200 and off by default): the quantised boxes contain
    the float boxes (emul_bvh_check codes 8-10), so the hierarchy still only culls and every output is identical."""
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 2)]
    outs = []
    for defs in ((), ("-DRT_NODEQ=1",)):
        E = eb.EmulScene(sc, rt.FP32_FAST, 4, 2, defs=defs)
        assert E.bvh_check()[0] == 0
        outs.append(E.render(sc["camera"], sc["lights"], 160, 120))
        E.close()
    for a, b in zip(outs[0][:5], outs[1][:5]):
        assert np.array_equal(a, b)
    assert outs[0][5][2] == outs[1][5][2] and outs[0][5][2] > 0       # hit pixels


def test_plain_and_postponing_any_hit_loops_render_the_identical_frame(example_scene, monkeypatch):
    """The two any-hit loops (leaves tested where they are met / postponed in batches of a bounded size) answer every shadow ray
    alike, whatever the bound: replay built with -DRT_ANY_DEFER=0 against the default build with bounds 4, 16 and 128."""
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 2)]
    E = eb.EmulScene(sc, rt.FP32_FAST, 4, 2, defs=("-DRT_ANY_DEFER=0",))
    ref = E.render(sc["camera"], sc["lights"], 160, 120)
    E.close()
    for bound in ("4", "16", "128"):
        monkeypatch.setenv("RT_B200_ANY_DEFER", bound)
        E = eb.EmulScene(sc, rt.FP32_FAST, 4, 2)
        out = E.render(sc["camera"], sc["lights"], 160, 120)
        E.close()
        for a, b in zip(ref[:5], out[:5]):
            assert np.array_equal(a, b)
