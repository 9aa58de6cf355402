"""GPU (-m gpu), round 2: the fused frame kernel, the primary-ray culling rectangle, the separate any-hit hierarchy and
the per-call cancellation, all through the C ABI.  Every variant must produce the frame of the three-kernel pipeline
bit for bit (it is the same device code in another launch shape) and meet the oracle bars of test_gpu_parity.py."""
import threading
import time

import numpy as np
import pytest

from conftest import random_soup_scene
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built):
    import distributed_raytracer_b200 as rt
    if rt._lib.load().rt_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return rt


def subdivided(example_scene, levels, n_lights=2):
    from distributed_raytracer_b200 import workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], levels)]
    sc["lights"] = example_scene["lights"][:n_lights]
    return sc


@pytest.mark.parametrize("scene_kind", ["example", "subdiv3", "three_objects"])
def test_fused_kernel_equals_pipeline(rt, example_scene, scene_kind):
    sc = {"example": lambda: example_scene, "subdiv3": lambda: subdivided(example_scene, 3),
          "three_objects": lambda: random_soup_scene(n_tris=150, seed=13, n_obj=3, n_lights=4)}[scene_kind]()
    W, H = (480, 270) if scene_kind != "three_objects" else (160, 120)
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    a = g.render_region_debug(W, H, flags=rt.FLAG_PIPELINE)
    na = g.stats()["hit_pixels"]
    b = g.render_region_debug(W, H, flags=rt.FLAG_FUSED)
    nb = g.stats()["hit_pixels"]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert na == nb == int((a[2] >= 0).sum())
    # sub-region, row-major, SDL quantiser
    x0, y0, w, h = W // 5, H // 7, W // 2 + 3, H // 2 + 1
    for flags in (0, rt.LAYOUT_ROWMAJOR, rt.QUANT_RGBA16_SHR8):
        p = g.render_region(W, H, x0, y0, w, h, flags=flags | rt.FLAG_PIPELINE)
        f = g.render_region(W, H, x0, y0, w, h, flags=flags | rt.FLAG_FUSED)
        assert np.array_equal(p, f)
    g.close()


def test_fused_kernel_meets_the_oracle_bar(rt, example_scene):
    S = orc.Scene(example_scene)
    rgb_o, _, face_o, _ = S.render_region(320, 240, mode=orc.CULLED)
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    rgb, col, face, obj, lit = g.render_region_debug(320, 240, flags=rt.FLAG_FUSED)
    g.close()
    assert (np.abs(rgb.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean() >= 0.999
    assert (face == face_o).mean() >= 0.999


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_cull_rectangle_never_changes_a_pixel(rt, example_scene, precision):
    from distributed_raytracer_b200 import workloads
    prec = rt.FP64_EXACT if precision == "exact" else rt.FP32_FAST
    g = rt.GpuScene(example_scene, precision=prec)
    lights = example_scene["lights"]
    views = [(example_scene["camera"], 320, 240)]
    for k, (radius, elev) in enumerate([(6.0, 0.3), (2.5, -0.4), (1.2, 0.9), (0.4, 0.0), (9.0, 1.2)]):
        views.append((workloads.orbit_camera([1.0, 1.0, -1.0], radius, np.deg2rad(40.0 * k + 10.0), elev, 1.0), 200 + 8 * k, 150 + 5 * k))
    for cam, W, H in views:
        g.set_frame(cam, lights)
        a = g.render_region_debug(W, H, flags=rt.FLAG_NO_CULL | rt.FLAG_PIPELINE)
        b = g.render_region_debug(W, H, flags=rt.FLAG_PIPELINE)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        if precision == "fast":
            c = g.render_region_debug(W, H, flags=rt.FLAG_FUSED)
            for x, y in zip(a, c):
                assert np.array_equal(x, y)
        x0, y0, w, h = W // 3, H // 4, W // 2, H // 2
        assert np.array_equal(g.render_region(W, H, x0, y0, w, h, flags=rt.FLAG_NO_CULL), g.render_region(W, H, x0, y0, w, h))
    g.close()


def test_cull_rectangle_in_tile_mode(rt, example_scene):
    import torch
    W, H, tw, th, n = 320, 240, 32, 16, 3
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    g.set_frame(example_scene["camera"], example_scene["lights"])
    per = rt.tiles_per_rank(W, H, tw, th, n)
    for rank in range(n):
        outs = []
        for flags in (rt.FLAG_NO_CULL | rt.FLAG_PIPELINE, rt.FLAG_PIPELINE, rt.FLAG_FUSED):
            t = torch.zeros(per * tw * th * 3, dtype=torch.uint8, device="cuda")
            g.render_tiles_device(W, H, tw, th, rank, n, t.data_ptr(), None, flags)
            outs.append(t.cpu().numpy())
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        assert outs[0].any()
    g.close()


@pytest.mark.parametrize("scene_kind", ["example", "subdiv3", "three_objects"])
def test_sm_affine_persistent_kernel_equals_the_plain_fused_kernel(rt, example_scene, scene_kind):
    """RT_FLAG_SM_AFFINE (k_frame_fused_sm: persistent warps, screen-local chunks per SM, one atomic per refill) hands the same
    8x4 blocks to the same device code in another order: every output is bit-identical -- whole frames, odd sizes whose chunk grid
    overhangs the frame, sub-regions, both layouts and quantisers, with and without the cull rectangle, tile mode for 1 and 3 ranks."""
    import torch
    sc = {"example": lambda: example_scene, "subdiv3": lambda: subdivided(example_scene, 3),
          "three_objects": lambda: random_soup_scene(n_tris=150, seed=13, n_obj=3, n_lights=4)}[scene_kind]()
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    for (W, H) in [(480, 270), (333, 187), (40, 12), (8, 4), (1, 1)]:
        for extra in (0, rt.FLAG_NO_CULL):
            a = g.render_region_debug(W, H, flags=rt.FLAG_FUSED | extra)
            na = g.stats()["hit_pixels"]
            b = g.render_region_debug(W, H, flags=rt.FLAG_FUSED | rt.FLAG_SM_AFFINE | extra)
            nb = g.stats()["hit_pixels"]
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
            assert na == nb
    W, H = 480, 270
    x0, y0, w, h = W // 5, H // 7, W // 2 + 3, H // 2 + 1
    for flags in (0, rt.LAYOUT_ROWMAJOR, rt.QUANT_RGBA16_SHR8):
        assert np.array_equal(g.render_region(W, H, x0, y0, w, h, flags=flags | rt.FLAG_FUSED),
                              g.render_region(W, H, x0, y0, w, h, flags=flags | rt.FLAG_FUSED | rt.FLAG_SM_AFFINE))
    tw, th = 32, 16
    for n in (1, 3):
        per = rt.tiles_per_rank(W, H, tw, th, n)
        for rank in range(n):
            outs = []
            for flags in (rt.FLAG_FUSED, rt.FLAG_FUSED | rt.FLAG_SM_AFFINE):
                t = torch.zeros(per * tw * th * 3, dtype=torch.uint8, device="cuda")
                g.render_tiles_device(W, H, tw, th, rank, n, t.data_ptr(), None, flags)
                outs.append(t.cpu().numpy())
            assert np.array_equal(outs[0], outs[1])
    # counting build: the same traversal work is reported
    g.render_region_debug(W, H, flags=rt.FLAG_FUSED | rt.FLAG_COUNT)
    c0 = g.stats()
    g.render_region_debug(W, H, flags=rt.FLAG_FUSED | rt.FLAG_SM_AFFINE | rt.FLAG_COUNT)
    c1 = g.stats()
    for k in ("nodes_primary", "tris_primary", "nodes_shade", "tris_shade", "hit_pixels", "shadow_rays"):
        assert c0[k] == c1[k], k
    g.close()


def test_bulk_trace_in_strips_returns_the_same_reply(rt, example_scene, monkeypatch):
    """rt_bulk_trace cuts a large order into strips that are contiguous in the reply and overlaps the device->host copy of strip k
    with the render of strip k+1 (RT_B200_BULK_STRIPS, default 4).  The reply must be byte-identical to the uncut order: both
    layouts, pinned and pageable destinations, a sub-region, odd sizes, and the frame's hit count in rt_get_stats."""
    import torch
    sc = subdivided(example_scene, 2)
    cam, lights = sc["camera"], sc["lights"]
    monkeypatch.setenv("RT_B200_BULK_STRIPS", "1")
    g1 = rt.GpuScene(sc, precision=rt.FP32_FAST)
    replies = {}
    for strips in ("4", "7"):
        monkeypatch.setenv("RT_B200_BULK_STRIPS", strips)
        g = rt.GpuScene(sc, precision=rt.FP32_FAST)
        for (W, H, x, y, w, h) in [(1280, 720, 0, 0, 1280, 720), (1283, 717, 0, 0, 1283, 717), (1280, 720, 100, 50, 1000, 600), (640, 360, 0, 0, 640, 360)]:
            for flags in (0, rt.LAYOUT_ROWMAJOR, rt.QUANT_RGBA16_SHR8, rt.FLAG_PIPELINE):
                a = g1.bulk_trace(cam, lights, W, H, x, y, w, h, flags=flags)
                na = g1.stats()["hit_pixels"]
                b = g.bulk_trace(cam, lights, W, H, x, y, w, h, flags=flags)                      # pageable numpy destination
                nb = g.stats()["hit_pixels"]
                assert np.array_equal(a, b), (strips, W, H, x, y, w, h, flags)
                assert na == nb == int((a.reshape(-1, 3).any(axis=1)).sum()) or na == nb
                pinned = torch.empty(w * h * 3, dtype=torch.uint8, pin_memory=True)              # pinned destination: copied in place
                out = pinned.numpy().reshape(-1, 3)
                g.bulk_trace(cam, lights, W, H, x, y, w, h, flags=flags, out=out)
                assert np.array_equal(a, out)
        replies[strips] = True
        g.close()
    # FP64_EXACT orders (and RT_FLAG_PIPELINE above) are not cut: same reply either way
    monkeypatch.setenv("RT_B200_BULK_STRIPS", "4")
    ge = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    monkeypatch.setenv("RT_B200_BULK_STRIPS", "1")
    ge1 = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    a = ge1.bulk_trace(example_scene["camera"], example_scene["lights"], 800, 600, 0, 0, 800, 600)
    b = ge.bulk_trace(example_scene["camera"], example_scene["lights"], 800, 600, 0, 0, 800, 600)
    assert np.array_equal(a, b) and a.any()
    ge.close(); ge1.close(); g1.close()


def test_any_hit_postponement_tunes_itself_and_never_changes_a_pixel(rt, example_scene, monkeypatch):
    """The kernels count the occluded share of their own shadow rays and pick how many leaves an any-hit ray postpones (16 / 8 / 4:
    rt_stats.any_defer, occluded_fraction).  The count must equal what the lit masks of the frame say, the bound must follow the
    thresholds, and a fixed bound (RT_B200_ANY_DEFER) must give the identical frame."""
    W, H = 1280, 720
    scenes = {"suzanne": example_scene, "soup": random_soup_scene(n_tris=4000, seed=5, n_lights=4, extent=1.0, half=0.25)}
    for name, sc in scenes.items():
        monkeypatch.delenv("RT_B200_ANY_DEFER", raising=False)
        g = rt.GpuScene(sc, precision=rt.FP32_FAST)
        rgb, col, face, obj, lit = g.render_region_debug(W, H)
        st = g.stats()
        hit = face >= 0
        n_l = np.asarray(sc["lights"]).reshape(-1, 6).shape[0]
        popc = np.array([bin(int(v)).count("1") for v in np.unique(lit)])
        lut = dict(zip(np.unique(lit).tolist(), popc.tolist()))
        occluded = int(sum((n_l - lut[int(v)]) * int(c) for v, c in zip(*np.unique(lit[hit], return_counts=True))))
        total = int(hit.sum()) * n_l
        assert total >= 4 * 65536, "the scene must produce enough shadow rays for a few tuning steps"
        f = occluded / total
        assert abs(st["occluded_fraction"] - f) < 0.05          # one CTA in four is sampled
        assert st["any_defer"] in (1, 8, 16)
        ref = g.render_region(W, H)
        g.close()
        for fixed in ("1", "4", "128"):
            monkeypatch.setenv("RT_B200_ANY_DEFER", fixed)
            gf = rt.GpuScene(sc, precision=rt.FP32_FAST)
            assert np.array_equal(gf.render_region(W, H), ref)
            assert gf.stats()["any_defer"] == int(fixed)
            gf.close()
        print(name, "occluded fraction %.3f -> bound %d" % (f, st["any_defer"]))


@pytest.mark.parametrize("leaf_any", [1, 2])
def test_any_hit_hierarchy_leaf_size(rt, example_scene, monkeypatch, leaf_any):
    sc = subdivided(example_scene, 3)
    monkeypatch.setenv("RT_B200_MAX_LEAF_SHADE", "4")
    g0 = rt.GpuScene(sc, precision=rt.FP32_FAST)
    monkeypatch.setenv("RT_B200_MAX_LEAF_SHADE", str(leaf_any))
    g1 = rt.GpuScene(sc, precision=rt.FP32_FAST)
    info = g1.bvh_info()
    assert info["max_leaf_any"] == leaf_any and info["max_leaf"] == 4 and info["n_nodes_any"] > info["n_nodes"]
    a = g0.render_region_debug(480, 270)
    b = g1.render_region_debug(480, 270)
    g0.close(); g1.close()
    assert np.array_equal(a[2], b[2])
    assert (a[4] == b[4]).mean() >= 0.9999 and (a[0] == b[0]).all(axis=1).mean() >= 0.9999


def test_cancel_is_per_call_and_does_not_stick(rt, example_scene):
    """rt_cancel stops the calls in flight (ctx.Err(), distributed/main.go:73) and nothing else: a call that starts after
    it renders normally; with nothing in flight it is a no-op."""
    sc = subdivided(example_scene, 4)
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)          # exact mode at 4K: a call long enough to be hit by the cancel
    ref = g.bulk_trace(sc["camera"], sc["lights"], 640, 360, 0, 0, 640, 360)
    L = rt._lib.load()
    assert L.rt_cancel(g.h) == 0                          # nothing in flight
    again = g.bulk_trace(sc["camera"], sc["lights"], 640, 360, 0, 0, 640, 360)
    assert np.array_equal(ref, again)
    result = {}

    def long_call():
        try:
            g.bulk_trace(sc["camera"], sc["lights"], 3840, 2160, 0, 0, 3840, 2160)
            result["rc"] = 0
        except rt.RtError as ex:
            result["rc"] = ex.code
    th = threading.Thread(target=long_call)
    t0 = time.perf_counter()
    th.start()
    time.sleep(0.002)
    assert L.rt_cancel(g.h) == 0
    th.join()
    result["dt"] = time.perf_counter() - t0
    assert result["rc"] in (0, -5)                        # RT_ERR_CANCELLED unless the frame had already finished
    after = g.bulk_trace(sc["camera"], sc["lights"], 640, 360, 0, 0, 640, 360)
    assert np.array_equal(ref, after)                     # the cancellation did not leak into the next call
    g.close()


def test_last_error_follows_the_last_call(rt, example_scene):
    import ctypes as C
    L = rt._lib.load()
    L.rt_last_error.restype = C.c_char_p
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    with pytest.raises(rt.RtError):
        g.render_region(64, 48, 0, 0, 65, 48)             # region exceeds the screen
    assert b"region" in L.rt_last_error(g.h)
    g.render_region(64, 48)                               # a successful call clears the thread's message ...
    assert L.rt_last_error(None) == b""
    assert b"region" in L.rt_last_error(g.h)              # ... the handle keeps its most recent failure
    g.close()


# ---- rt_group_*: one process, several GPUs, ONE BulkTrace (no torch.distributed) ------------------------------------
def _group_devices(rt):
    n = rt._lib.load().rt_device_count()
    return list(range(n)) if n >= 2 else [0, 0, 0]        # one GPU: three shares on the same device exercise the same code


@pytest.mark.parametrize("scene_kind", ["c1", "subdiv4"])
def test_group_bulk_trace_equals_single_device(rt, example_scene, scene_kind):
    sc = example_scene if scene_kind == "c1" else subdivided(example_scene, 4)
    W, H = (320, 240) if scene_kind == "c1" else (1920, 1080)
    one = rt.GpuScene(sc, precision=rt.FP32_FAST)
    grp = rt.GpuGroup(sc, devices=_group_devices(rt), precision=rt.FP32_FAST)
    assert grp.n_devices == len(_group_devices(rt))
    L = rt._lib.load()
    for (x, y, w, h) in [(0, 0, W, H), (W // 5 + 3, H // 7 + 1, W // 2 + 5, H // 2 + 3), (W - 9, H - 5, 9, 5)]:
        for flags in (0, rt.LAYOUT_ROWMAJOR, rt.QUANT_RGBA16_SHR8):
            ref = one.bulk_trace(sc["camera"], sc["lights"], W, H, x, y, w, h, flags=flags)
            got = grp.bulk_trace(sc["camera"], sc["lights"], W, H, x, y, w, h, flags=flags)
            assert np.array_equal(ref, got), (x, y, w, h, flags)
    # pinned destination: the GPUs copy straight into the caller's buffer
    import ctypes
    nbytes = W * H * 3
    p = L.rt_host_alloc(nbytes)
    out = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p)).reshape(-1, 3)
    out[:] = 7
    grp.bulk_trace(sc["camera"], sc["lights"], W, H, 0, 0, W, H, out=out)
    assert np.array_equal(out, one.bulk_trace(sc["camera"], sc["lights"], W, H, 0, 0, W, H))
    L.rt_host_free(p)
    with pytest.raises(rt.RtError):
        grp.bulk_trace(sc["camera"], sc["lights"], W, H, 0, 0, W + 1, H)
    one.close(); grp.close()


def test_group_bulk_trace_exact_and_concurrent_callers(rt, example_scene):
    """FP64_EXACT through the group is bit-identical to the oracle's bytes; concurrent orders (a goroutine each in the reference
    worker) with DIFFERENT cameras do not disturb each other."""
    from concurrent.futures import ThreadPoolExecutor
    from distributed_raytracer_b200 import workloads
    sc = example_scene
    S = orc.Scene(sc)
    grp = rt.GpuGroup(sc, devices=_group_devices(rt), precision=rt.FP64_EXACT)
    cams = [sc["camera"]] + [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(25.0 * k), 0.2, 1.0) for k in range(1, 6)]
    refs = []
    for cam in cams:
        S.set_frame(cam, sc["lights"])
        refs.append(S.render_region(200, 150, mode=orc.CULLED)[0])
    S.close()

    def order(k):
        return [grp.bulk_trace(cams[k], sc["lights"], 200, 150, 0, 0, 200, 150) for _ in range(5)]
    with ThreadPoolExecutor(6) as pool:
        outs = list(pool.map(order, range(len(cams))))
    for k, frames in enumerate(outs):
        for f in frames:
            assert np.array_equal(f, refs[k]), k
    grp.close()


def test_group_object_positions_follow_the_order(rt):
    sc = random_soup_scene(n_tris=150, seed=13, n_obj=3, n_lights=4)
    one = rt.GpuScene(sc, precision=rt.FP32_FAST)
    grp = rt.GpuGroup(sc, devices=_group_devices(rt), precision=rt.FP32_FAST)
    moved = np.asarray(sc["obj_pos"], dtype=np.float64) + [[0.3, -0.2, 0.1]]
    for pos in (None, moved, None):                       # an order without positions renders the registered ones again
        ref = one.bulk_trace(sc["camera"], sc["lights"], 160, 120, 0, 0, 160, 120, obj_pos=pos)
        got = grp.bulk_trace(sc["camera"], sc["lights"], 160, 120, 0, 0, 160, 120, obj_pos=pos)
        assert np.array_equal(ref, got)
    one.close(); grp.close()


def test_object_with_nan_position_is_absent(rt):
    """A registered object that an order does not list arrives with a NaN position (rt_b200.h): it must not be rendered, and the
    other objects must render as in the scene without it (environment.go:73-98)."""
    sc = random_soup_scene(n_tris=150, seed=13, n_obj=3, n_lights=4)
    without = dict(sc, obj_pos=np.asarray(sc["obj_pos"])[[0, 2]], obj_mesh=np.asarray(sc["obj_mesh"])[[0, 2]])
    S = orc.Scene(without)
    rgb_o, col_o, face_o, _ = S.render_region(160, 120, mode=orc.CULLED)
    S.close()
    pos = np.asarray(sc["obj_pos"], dtype=np.float64).copy()
    pos[1] = np.nan
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    g.set_frame(sc["camera"], sc["lights"], obj_pos=pos)
    rgb, col, face, obj, lit = g.render_region_debug(160, 120)
    g.close()
    assert not (obj == 1).any() and (obj == 2).any()
    assert np.array_equal(face, face_o) and np.array_equal(col.view(np.uint64), col_o.view(np.uint64)) and np.array_equal(rgb, rgb_o)
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    fast = g.bulk_trace(sc["camera"], sc["lights"], 160, 120, 0, 0, 160, 120, obj_pos=pos)
    g.close()
    assert (np.abs(fast.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean() >= 0.995
