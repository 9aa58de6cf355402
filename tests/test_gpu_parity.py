"""GPU (-m gpu): parity of the sm_100a kernels with the oracle, called through the C ABI (librt_b200.so).

Bars (BASELINE.json north_star): FP64_EXACT -- nearest-hit face index per primary ray and the float64 colours
bit-exact, bytes identical under both quantisers; FP32_FAST -- <= 1 LSB per 8-bit channel on >= 99.9 % of the
pixels and the hit-face index equal on >= 99.9 %."""
import numpy as np
import pytest

from conftest import degenerate_scenes, EXAMPLE_SCENE, random_soup_scene
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

LSB_TOL = 1          # FP32_FAST: max |delta| per 8-bit channel ...
LSB_FRAC = 0.999     # ... on at least this fraction of the pixels


def tile_owner(i, j, W, tw, th, n):
    """csrc/rt_types.h tile_id(): tile row Y is rotated by 3*Y positions; owner = id % n_ranks."""
    tiles_x = (W + tw - 1) // tw
    ty = j // th
    return (ty * tiles_x + (i // tw - 3 * ty) % tiles_x) % n


@pytest.fixture(scope="module")
def rt(built):
    import distributed_raytracer_b200 as rt
    if rt._lib.load().rt_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return rt


def oracle_render(sc, W, H, x=0, y=0, w=None, h=None, **kw):
    S = orc.Scene(sc)
    return S.render_region(W, H, x, y, w, h, mode=orc.CULLED, **kw)


def assert_exact(rt, sc, W, H, flags=0):
    rgb_o, col_o, face_o, st = oracle_render(sc, W, H)
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    rgb, col, face, obj, lit = g.render_region_debug(W, H, flags=flags)
    stats = g.stats()
    g.close()
    assert np.array_equal(face, face_o), "%d face mismatches" % (face != face_o).sum()
    assert np.array_equal(col.view(np.uint64), col_o.view(np.uint64))
    assert np.array_equal(rgb, rgb_o)
    assert stats["hit_pixels"] == st["hit_pixels"] and stats["shadow_rays"] == st["shadow_rays"]
    return st


def fast_fractions(rt, sc, W, H, flags=0):
    rgb_o, col_o, face_o, st = oracle_render(sc, W, H)
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    rgb, col, face, obj, lit = g.render_region_debug(W, H, flags=flags)
    g.close()
    within = (np.abs(rgb.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= LSB_TOL).mean()
    return within, (face == face_o).mean(), rgb, face


def test_c1_exact_bit_identical(rt, example_scene):
    st = assert_exact(rt, example_scene, 320, 240)
    assert st["hit_pixels"] == 5820


def test_c1_exact_matches_golden_fixture(rt, example_scene):
    """GPU FP64_EXACT against the golden frame of the INDEPENDENT restatement (tests/golden/make_golden_pyref.py), every pixel:
    face map, float64 colour bits, lit masks, both quantisers."""
    import os
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "c1_320x240_pyref.npz"))
    g = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    rgb, col, face, obj, lit = g.render_region_debug(320, 240)
    sdl = g.render_region(320, 240, flags=rt.QUANT_RGBA16_SHR8)
    g.close()
    hit = gold["hit_index"]
    assert np.array_equal(face, gold["face"].astype(np.int32))
    assert np.array_equal(col[hit].view(np.uint64), gold["colour_bits"])
    assert np.array_equal(lit[hit], gold["lit"].astype(np.uint32))
    assert np.array_equal(rgb[hit], gold["rgb"]) and np.array_equal(sdl[hit], gold["sdl"])
    assert not rgb[face < 0].any() and (obj[face >= 0] == 0).all() and (obj[face < 0] == -1).all()


def test_c2_exact_matches_golden_sample(rt, example_scene):
    """C2 (1920x1080, lights[0]): every 97th pixel of the GPU FP64_EXACT frame against the independent restatement; FP32_FAST on
    the same sample: same face on >= 99.9 %, within 1 LSB on >= 99.9 %."""
    import os
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "c2_1080p_sample_pyref.npz"))
    sc = dict(example_scene)
    sc["lights"] = example_scene["lights"][:1]
    idx = gold["index"]
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    rgb, col, face, obj, lit = g.render_region_debug(1920, 1080)
    g.close()
    assert np.array_equal(face[idx], gold["face"].astype(np.int32))
    assert np.array_equal(col[idx].view(np.uint64), gold["colour_bits"])
    assert np.array_equal(lit[idx], gold["lit"].astype(np.uint32))
    assert np.array_equal(rgb[idx], gold["rgb"])
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    rgb, col, face, obj, lit = g.render_region_debug(1920, 1080)
    g.close()
    assert (face[idx] == gold["face"]).mean() >= LSB_FRAC
    assert (np.abs(rgb[idx].astype(np.int32) - gold["rgb"].astype(np.int32)).max(axis=1) <= LSB_TOL).mean() >= LSB_FRAC


def test_c1_exact_brute_force_kernel(rt, example_scene):
    assert_exact(rt, example_scene, 160, 120, flags=rt.FLAG_BRUTE_FORCE)


def test_c2_exact_1080p_one_light(rt, example_scene):
    sc = dict(example_scene)
    sc["lights"] = example_scene["lights"][:1]            # C2: lights[0] only
    assert_exact(rt, sc, 1920, 1080)


@pytest.mark.parametrize("kw", [dict(n_tris=300, seed=11), dict(n_tris=500, seed=12, with_normals=True),
                                dict(n_tris=150, seed=13, n_obj=3, n_lights=4), dict(n_tris=1, seed=14, half=1.2, extent=0.2),
                                dict(n_tris=2, seed=15, half=1.2, extent=0.2), dict(n_tris=40, seed=17, half=1.5)])
def test_exact_random_scenes(rt, kw):
    st = assert_exact(rt, random_soup_scene(**kw), 160, 120)
    assert st["hit_pixels"] > 0


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_soup_flat_normals_eight_lights_orbit_camera(rt, example_scene, precision):
    """Config-4-shaped scene at test size (random soup without vertex normals => Triangle.Normal path, default material, 8 white
    lights on a ring) seen from a rotated orbit camera (config 5)."""
    from test_emul import soup_config_scene
    sc = soup_config_scene(20000, example_scene, angle_deg=33.0)
    if precision == "exact":
        st = assert_exact(rt, sc, 320, 180)
        assert st["hit_pixels"] > 2000 and st["shadow_rays"] == 8 * st["hit_pixels"]
    else:
        within, same, _, _ = fast_fractions(rt, sc, 320, 180)
        assert within >= 0.995 and same >= 0.995      # a soup of tiny triangles is all silhouette


def test_exact_duplicate_triangles_tie_break(rt):
    sc = random_soup_scene(n_tris=20, seed=21)
    m = sc["meshes"][0]
    m["verts"] = np.concatenate([m["verts"]] * 4, axis=0)
    f = m["faces"].copy()
    m["faces"] = np.concatenate([f + [60 * k, 60 * k, 60 * k, 0, 0, 0, 0] for k in range(4)]).astype(np.uint32)
    _, _, face_o, st = oracle_render(sc, 64, 48)
    assert st["nearest_ties"] > 0
    assert_exact(rt, sc, 64, 48)


def test_exact_subdivided_suzanne(rt, example_scene):
    from distributed_raytracer_b200 import workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 3)]      # 61 952 faces
    sc["lights"] = example_scene["lights"][:2]
    assert_exact(rt, sc, 480, 270)


def test_exact_region_layout_quantiser(rt, example_scene):
    W, H, x, y, w, h = 203, 151, 61, 40, 83, 67          # odd sizes: integer halves (tracer.go:16), ragged blocks
    g = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    for flags, layout, quant in [(0, orc.L_COLMAJOR, orc.Q_RGB255), (rt.LAYOUT_ROWMAJOR, orc.L_ROWMAJOR, orc.Q_RGB255),
                                 (rt.QUANT_RGBA16_SHR8, orc.L_COLMAJOR, orc.Q_RGBA16_SHR8),
                                 (rt.LAYOUT_ROWMAJOR | rt.QUANT_RGBA16_SHR8, orc.L_ROWMAJOR, orc.Q_RGBA16_SHR8)]:
        rgb_o, _, face_o, _ = oracle_render(example_scene, W, H, x, y, w, h, layout=layout, quant=quant)
        rgb = g.render_region(W, H, x, y, w, h, flags=flags)
        assert np.array_equal(rgb, rgb_o)
    g.close()


def test_fast_c1_within_tolerance_both_kernels(rt, example_scene):
    w1, f1, rgb_p, face_p = fast_fractions(rt, example_scene, 320, 240, flags=rt.FLAG_PERSISTENT)
    w2, f2, rgb_s, face_s = fast_fractions(rt, example_scene, 320, 240, flags=rt.FLAG_SIMPLE)
    assert w1 >= LSB_FRAC and f1 >= LSB_FRAC and w2 >= LSB_FRAC and f2 >= LSB_FRAC
    # both kernels run the same per-ray arithmetic => identical images
    assert np.array_equal(face_p, face_s) and np.array_equal(rgb_p, rgb_s)


def test_fast_1080p_two_lights(rt, example_scene):
    sc = dict(example_scene)
    sc["lights"] = example_scene["lights"][:2]
    within, same, _, _ = fast_fractions(rt, sc, 1920, 1080)
    assert within >= LSB_FRAC and same >= LSB_FRAC


def test_fast_subdivided_suzanne(rt, example_scene):
    from distributed_raytracer_b200 import workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 3)]
    sc["lights"] = example_scene["lights"][:2]
    within, same, _, _ = fast_fractions(rt, sc, 960, 540)
    assert within >= LSB_FRAC and same >= LSB_FRAC


@pytest.mark.parametrize("kw", [dict(n_tris=400, seed=31, with_normals=True), dict(n_tris=150, seed=32, n_obj=3, n_lights=4),
                                dict(n_tris=1, seed=14, half=1.2, extent=0.2)])
def test_fast_random_scenes(rt, kw):
    within, same, _, _ = fast_fractions(rt, random_soup_scene(**kw), 160, 120)
    assert within >= 0.995 and same >= 0.995          # tiny soups: many silhouette pixels per frame


@pytest.mark.parametrize("W,H,tw,th,n,layout", [(333, 187, 32, 16, 3, "row"), (352, 192, 32, 16, 3, "row"), (352, 192, 32, 16, 8, "col"),
                                               (340, 200, 64, 32, 2, "col"), (333, 187, 8, 4, 5, "col")])
def test_fast_tiles_equal_region_render(rt, example_scene, W, H, tw, th, n, layout):
    """Multi-GPU tile path on one GPU: every rank's packed tiles, gathered and de-interleaved (fast word path for interior
    tiles, byte path for tiles cut by the screen edge), reproduce the plain region render byte for byte."""
    import torch
    flags = rt.LAYOUT_ROWMAJOR if layout == "row" else rt.LAYOUT_COLMAJOR
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H, flags=flags)
    per = rt.tiles_per_rank(W, H, tw, th, n)
    allt = torch.zeros(n * per * tw * th * 3, dtype=torch.uint8, device="cuda")
    for r in range(n):
        part = allt[r * per * tw * th * 3:(r + 1) * per * tw * th * 3]
        g.render_tiles_device(W, H, tw, th, r, n, part.data_ptr(), None)
    out = torch.full((W * H * 3,), 77, dtype=torch.uint8, device="cuda")
    g.untile_device(W, H, tw, th, n, allt.data_ptr(), out.data_ptr(), None, flags=flags)
    torch.cuda.synchronize()
    g.close()
    assert np.array_equal(out.cpu().numpy().reshape(-1, 3), ref)


@pytest.mark.parametrize("W,H,tw,th,n,layout", [(352, 192, 32, 16, 3, "col"), (352, 192, 32, 16, 8, "row"), (333, 187, 32, 16, 4, "col")])
def test_fast_peer_writer_equals_region_render(rt, example_scene, W, H, tw, th, n, layout):
    """rt_render_tiles_peer (the framebuffer writer storing straight into the full framebuffer, as it does over NVLink into rank
    0's memory): all ranks emulated on one GPU, writing into ONE shared buffer through the library's shareable allocation,
    must reproduce the region render; bytes of other ranks' tiles must never be touched."""
    import torch
    flags = rt.LAYOUT_ROWMAJOR if layout == "row" else rt.LAYOUT_COLMAJOR
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H, flags=flags)
    buf = rt.DeviceBuffer(0, W * H * 3)
    assert len(buf.export()) == 64
    fb = torch.as_tensor(buf, device="cuda")
    fb.fill_(77)
    g.render_tiles_peer(W, H, tw, th, 1, n, buf.ptr, None, flags)           # one rank only: the rest must keep the fill value
    one = fb.cpu().numpy().reshape(-1, 3).copy()
    i, j = (np.arange(W * H) % W, np.arange(W * H) // W) if layout == "row" else (np.arange(W * H) // H, np.arange(W * H) % H)
    owner = tile_owner(i, j, W, tw, th, n)
    assert (one[owner != 1] == 77).all() and np.array_equal(one[owner == 1], ref[owner == 1])
    for r in range(n):
        g.render_tiles_peer(W, H, tw, th, r, n, buf.ptr, None, flags)
    torch.cuda.synchronize()
    assert np.array_equal(fb.cpu().numpy().reshape(-1, 3), ref)
    g.close()
    del fb
    buf.close()


@pytest.mark.parametrize("W,H,tw,th,n,layout", [(352, 192, 32, 16, 3, "col"), (352, 192, 32, 16, 8, "row"), (333, 187, 32, 16, 4, "col"),
                                               (3840, 2160, 32, 16, 8, "col"), (640, 360, 8, 4, 2, "row")])
def test_fast_push_tiles_equals_region_render(rt, example_scene, W, H, tw, th, n, layout):
    """rt_render_tiles_device + rt_push_tiles_device (what every rank does towards rank 0's framebuffer over NVLink): all ranks
    emulated on one GPU pushing into ONE shareable buffer reproduce the region render; a single rank leaves every other
    rank's bytes untouched (16-byte vector path for 4K, word and byte paths for the odd sizes)."""
    import torch
    flags = rt.LAYOUT_ROWMAJOR if layout == "row" else rt.LAYOUT_COLMAJOR
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H, flags=flags)
    buf = rt.DeviceBuffer(0, W * H * 3)
    fb = torch.as_tensor(buf, device="cuda")
    fb.fill_(77)
    per = rt.tiles_per_rank(W, H, tw, th, n)
    tiles = torch.zeros(per * tw * th * 3, dtype=torch.uint8, device="cuda")
    i, j = (np.arange(W * H) % W, np.arange(W * H) // W) if layout == "row" else (np.arange(W * H) // H, np.arange(W * H) % H)
    owner = tile_owner(i, j, W, tw, th, n)
    for r in [1] + [q for q in range(n) if q != 1]:
        g.render_tiles_device(W, H, tw, th, r, n, tiles.data_ptr(), None)
        g.push_tiles_device(W, H, tw, th, r, n, tiles.data_ptr(), buf.ptr, None, flags)
        if r == 1:
            one = fb.cpu().numpy().reshape(-1, 3)
            assert (one[owner != 1] == 77).all() and np.array_equal(one[owner == 1], ref[owner == 1])
    assert np.array_equal(fb.cpu().numpy().reshape(-1, 3), ref)
    g.close()
    del fb
    buf.close()


def test_push_tiles_into_host_shared_framebuffer(rt, example_scene):
    """rt_host_register: the framebuffer and its hand-off record live in pinned POSIX shared memory; every "rank" stores its
    tiles straight into host memory (second mapping of the same segment = what another process sees), flags included."""
    import torch
    W, H, tw, th, n = 352, 192, 32, 16, 3
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H)
    fb, ctrl = rt.HostSharedBuffer(0, W * H * 3), rt.HostSharedBuffer(0, 256)
    fb2 = rt.HostSharedBuffer(0, name=fb.name)                       # as a peer process would attach
    per = rt.tiles_per_rank(W, H, tw, th, n)
    tiles = torch.zeros(per * tw * th * 3, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    fs = rt.FrameState(example_scene["camera"], example_scene["lights"])
    for r in range(n):
        g.render_push_tiles(fs, W, H, tw, th, r, n, tiles.data_ptr(), (fb if r else fb2).ptr, st.cuda_stream, rt.LAYOUT_COLMAJOR, ctrl.ptr, 1)
    g.frame_complete_device(ctrl.ptr, 1, n, True, st.cuda_stream)
    st.synchronize()
    words = ctrl.array.view(np.int32)
    assert list(words[:n]) == [1] * n and words[32] == 1 and words[33] == 0
    assert np.array_equal(fb.array[:W * H * 3].reshape(-1, 3), ref)
    g.close()
    for b in (fb2, ctrl, fb):
        b.close()


@pytest.mark.parametrize("W,H,stripe,n,layout", [(352, 192, 8, 3, "col"), (333, 187, 16, 4, "col"), (352, 190, 8, 8, "row"), (3840, 2160, 8, 8, "col")])
def test_render_stripes_host_equals_region_render(rt, example_scene, W, H, stripe, n, layout):
    """rt_render_stripes_host (the multi-GPU end-to-end path): every "rank" renders its stripes in place and moves them with one
    strided D2H transfer into the pinned shared host framebuffer; all ranks together reproduce the region render, a single
    rank leaves the other ranks' bytes untouched (cut last stripe, H not a multiple of 4, both layouts, 4K)."""
    import torch
    flags = rt.LAYOUT_ROWMAJOR if layout == "row" else rt.LAYOUT_COLMAJOR
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H, flags=flags)
    fb, ctrl = rt.HostSharedBuffer(0, W * H * 3), rt.HostSharedBuffer(0, 256)
    fb.array[:] = 77
    local = torch.full((W * H * 3,), 55, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    fs = rt.FrameState(example_scene["camera"], example_scene["lights"])
    pix = np.arange(W * H)
    along = (pix // W if layout == "row" else pix // H)               # row index (row-major) / column index (column-major)
    owner = (along // stripe) % n
    for r in [1] + [q for q in range(n) if q != 1]:
        g.render_stripes_host(fs, W, H, stripe, r, n, local.data_ptr(), fb.ptr, st.cuda_stream, flags, ctrl.ptr, 1)
        st.synchronize()
        if r == 1:
            one = fb.array[:W * H * 3].reshape(-1, 3)
            assert (one[owner != 1] == 77).all() and np.array_equal(one[owner == 1], ref[owner == 1])
    g.frame_complete_device(ctrl.ptr, 1, n, True, st.cuda_stream)
    st.synchronize()
    words = ctrl.array.view(np.int32)
    assert list(words[:n]) == [1] * n and words[32] == 1 and words[33] == 0
    assert np.array_equal(fb.array[:W * H * 3].reshape(-1, 3), ref)
    g.close()
    for b in (ctrl, fb):
        b.close()


def test_frame_handoff_flags_two_ranks_on_one_gpu(rt, example_scene):
    """Device-side frame hand-off (rt_render_push_tiles with a sync record + rt_frame_complete_device): two "ranks" (two scene
    handles, two streams) on one GPU push into one framebuffer slot.  Rank 0's complete kernel is enqueued BEFORE rank 1 has
    even been asked to render, so it really waits on the device for rank 1's arrive flag; rank 1's next frame really waits for
    the slot to be released (the release is enqueued after a consumer copy).  Every consumed frame must equal the region render."""
    import torch
    from distributed_raytracer_b200 import workloads
    W, H, tw, th, n = 352, 192, 32, 16, 2
    flags = rt.LAYOUT_COLMAJOR
    scenes = [rt.GpuScene(example_scene, precision=rt.FP32_FAST) for _ in range(n)]
    streams = [torch.cuda.Stream() for _ in range(n)]
    fbuf, ctrl = rt.DeviceBuffer(0, W * H * 3), rt.DeviceBuffer(0, 256)
    fb = torch.as_tensor(fbuf, device="cuda")
    cw = torch.as_tensor(ctrl, device="cuda").view(torch.int32)
    fb.zero_(); cw.zero_()
    per = rt.tiles_per_rank(W, H, tw, th, n)
    tiles = [torch.zeros(per * tw * th * 3, dtype=torch.uint8, device="cuda") for _ in range(n)]
    cams = [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(20.0 * f), 0.05 * f, 1.04719755) for f in range(5)]
    consumed = [torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda") for _ in cams]
    for r in range(n):      # first use of a lane allocates its scratch (cudaMalloc waits for the device): not under a spinning kernel
        scenes[r].render_tiles_device(W, H, tw, th, r, n, tiles[r].data_ptr(), streams[r].cuda_stream)
    torch.cuda.synchronize()
    for f, cam in enumerate(cams):
        fs = rt.FrameState(cam, example_scene["lights"])
        seq = f + 1
        scenes[0].render_push_tiles(fs, W, H, tw, th, 0, n, tiles[0].data_ptr(), fbuf.ptr, streams[0].cuda_stream, flags, ctrl.ptr, seq)
        scenes[0].frame_complete_device(ctrl.ptr, seq, n, False, streams[0].cuda_stream)
        # Both "ranks" share ONE device here: nothing may be queued behind the waiting kernel in its own stream before the
        # other rank's work has been submitted (streams can share a hardware queue, and a queue stalls at the first entry
        # that must wait for the running kernel) -- across real GPUs the ranks' queues are independent.
        scenes[1].render_push_tiles(fs, W, H, tw, th, 1, n, tiles[1].data_ptr(), fbuf.ptr, streams[1].cuda_stream, flags, ctrl.ptr, seq)
        with torch.cuda.stream(streams[0]):
            torch.add(fb, 0, out=consumed[f])                            # the consumer
        scenes[0].frame_release_device(ctrl.ptr, seq, streams[0].cuda_stream)
    torch.cuda.synchronize()
    c = cw.cpu().numpy()
    assert list(c[:2]) == [5, 5] and c[32] == 5 and c[33] == 0, c[:34]
    ref = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    for f, cam in enumerate(cams):
        ref.set_frame(cam, example_scene["lights"])
        assert np.array_equal(consumed[f].cpu().numpy().reshape(-1, 3), ref.render_region(W, H, flags=flags)), f
    for g in scenes + [ref]:
        g.close()
    del fb, cw
    fbuf.close(); ctrl.close()


def test_frame_handoff_times_out_instead_of_hanging(rt, example_scene, monkeypatch):
    """A rank that never arrives must not hang the GPU: the wait gives up after RT_B200_SYNC_TIMEOUT_MS and records 0x100+rank."""
    import torch
    monkeypatch.setenv("RT_B200_SYNC_TIMEOUT_MS", "50")
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ctrl = rt.DeviceBuffer(0, 256)
    cw = torch.as_tensor(ctrl, device="cuda").view(torch.int32)
    cw.zero_()
    cw[0] = 1                                                            # rank 0 arrived, rank 1 never does
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    g.frame_complete_device(ctrl.ptr, 1, 2, True, st.cuda_stream)
    st.synchronize()
    c = cw.cpu().numpy()
    assert c[33] == 0x101 and c[32] == 1
    g.close()
    del cw
    ctrl.close()


def test_fast_device_output_and_streams(rt, example_scene):
    import torch
    W, H = 320, 240
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    ref = g.render_region(W, H)
    out = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g.render_region_device(W, H, 0, 0, W, H, out.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy().reshape(-1, 3), ref)
    out.zero_()
    g.render_region_device(W, H, 0, 0, W, H, out.data_ptr(), None)          # library stream, synchronous
    assert np.array_equal(out.cpu().numpy().reshape(-1, 3), ref)
    g.close()


def test_full_size_c3_properties(rt, example_scene):
    """BASELINE config 3 at full size (991 232 faces, 3840x2160, 2 lights): size-independent properties, then the oracle on
    EVERY pixel of the frame -- FP32_FAST within the north_star tolerance, FP64_EXACT bit-identical."""
    from distributed_raytracer_b200 import workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 5)]
    assert sc["meshes"][0]["faces"].shape[0] == 991232
    sc["lights"] = example_scene["lights"][:2]
    W, H = 3840, 2160
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    info = g.bvh_info()
    assert info["n_faces"] == 991232 and 0 < info["n_nodes"] < 991232 and info["max_depth"] < 94
    rgb, col, face, obj, lit = g.render_region_debug(W, H)
    stats = g.stats()
    # determinism + the two primary kernels agree (persistent refill order must not change any pixel)
    rgb2 = g.render_region(W, H)
    rgb3, _, face3, _, _ = g.render_region_debug(W, H, flags=rt.FLAG_PERSISTENT)
    assert np.array_equal(rgb, rgb2) and np.array_equal(rgb, rgb3) and np.array_equal(face, face3)
    hit = face >= 0
    assert stats["hit_pixels"] == int(hit.sum()) and stats["primary_rays"] == W * H
    assert 0.08 < hit.mean() < 0.12                                       # ~10.1 % of the primary rays hit
    assert not rgb[~hit].any() and (face[hit] < 991232).all()
    # subdividing does not move the surface: a face index maps back to its coarse parent (f >> 10), so the
    # coarse render of the original 968-face mesh must agree on the parent index almost everywhere
    g.close()
    g0 = rt.GpuScene(dict(example_scene, lights=sc["lights"]), precision=rt.FP32_FAST)
    _, _, face0, _, _ = g0.render_region_debug(W, H)
    g0.close()
    both = hit & (face0 >= 0)
    assert ((face >= 0) == (face0 >= 0)).mean() > 0.9995
    assert ((face[both] >> 10) == face0[both]).mean() > 0.999
    # the oracle on the WHOLE 4K frame (8.3 M pixels, 9.97 M rays: ~2 s on the box's host threads in its culling-BVH mode):
    # FP32_FAST tolerance of north_star on every pixel ...
    S = orc.Scene(sc)
    rgb_o, col_o, face_o, st = S.render_region(W, H, mode=orc.CULLED, n_threads=0)
    S.close()
    assert st["hit_pixels"] == 838336 and st["shadow_rays"] == 2 * 838336
    within = (np.abs(rgb.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= LSB_TOL).mean()
    same = (face == face_o).mean()
    assert within >= LSB_FRAC and same >= LSB_FRAC, (within, same)
    # ... and FP64_EXACT bit-identical on every pixel: face map, float64 colour bits, bytes
    ge = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    rgb_e, col_e, face_e, _, _ = ge.render_region_debug(W, H)
    ge.close()
    assert np.array_equal(face_e, face_o)
    assert np.array_equal(col_e.view(np.uint64), col_o.view(np.uint64))
    assert np.array_equal(rgb_e, rgb_o)


def test_c3_matches_golden_sample_of_the_independent_restatement(rt, example_scene):
    """Config 3 on the GPU against tests/golden/c3_4k_sample_pyref.npz (404 pixels by the independent restatement, incl. its own
    subdivision): FP64_EXACT bit-identical (face, float64 colour bits, lit mask, bytes), FP32_FAST same face / within 1 LSB on
    >= 99 % of this small sample."""
    import os
    from conftest import GOLDEN
    from distributed_raytracer_b200 import workloads
    gold = np.load(os.path.join(GOLDEN, "c3_4k_sample_pyref.npz"))
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 5)]
    sc["lights"] = example_scene["lights"][:2]
    W, H = 3840, 2160
    idx = gold["pixel_i"].astype(np.int64) * H + gold["pixel_j"].astype(np.int64)          # BulkTrace order
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    rgb, col, face, obj, lit = g.render_region_debug(W, H)
    g.close()
    hit = gold["face"] >= 0
    assert np.array_equal(face[idx], gold["face"])
    assert np.array_equal(col[idx][hit].view(np.uint64), gold["colour_bits"][hit])
    assert np.array_equal(lit[idx][hit], gold["lit"][hit].astype(np.uint32))
    assert np.array_equal(rgb[idx], gold["rgb"])
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    rgb, col, face, obj, lit = g.render_region_debug(W, H)
    g.close()
    assert (face[idx] == gold["face"]).mean() >= 0.99
    assert (np.abs(rgb[idx].astype(np.int32) - gold["rgb"].astype(np.int32)).max(axis=1) <= LSB_TOL).mean() >= 0.99


def test_counting_build_reports_traversal_work(rt, example_scene):
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    a = g.render_region(320, 240)
    s0 = g.stats()
    b = g.render_region(320, 240, flags=rt.FLAG_COUNT)
    s1 = g.stats()
    g.close()
    assert np.array_equal(a, b)
    assert s0["nodes_fetched"] == 0 and s1["nodes_fetched"] > 0 and s1["tris_fetched"] > 0
    assert s1["node_bytes"] == 112 and s1["tri_bytes"] == 64 and s1["primary_rays"] == 320 * 240 and s1["ms_total"] > 0


def test_errors_and_lifecycle(rt, example_scene):
    sc = {k: v for k, v in example_scene.items() if k not in ("camera", "lights")}
    g = rt.GpuScene(sc)
    with pytest.raises(rt.RtError) as ei:
        g.render_region(64, 48)                                 # render before rt_frame_set
    assert ei.value.code == rt._lib.RT_ERR_STATE
    g.set_frame(example_scene["camera"], example_scene["lights"])
    with pytest.raises(rt.RtError) as ei:
        g.render_region(64, 48, 32, 0, 64, 48)                  # region exceeds the screen
    assert ei.value.code == rt._lib.RT_ERR_INVALID
    ref = g.render_region(64, 48)
    g.set_frame(example_scene["camera"], example_scene["lights"][:0])      # no lights: ambient only
    amb = g.render_region(64, 48)
    assert set(np.unique(amb)) <= {0, int(255 * float(np.float32(0.1)))}
    g.close()
    # re-registration = destroy + create, repeatedly (distributed/main.go:146-185): must be leak-free and repeatable
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(5):
        g = rt.GpuScene(example_scene)
        assert np.array_equal(g.render_region(64, 48), ref)
        g.close()
    assert abs(torch.cuda.mem_get_info()[0] - free0) < 64 << 20
    # object positions are per-frame state (the gob diff): moving the object moves the image
    g = rt.GpuScene(example_scene)
    moved = g.bulk_trace(example_scene["camera"], example_scene["lights"], 64, 48, 0, 0, 64, 48, obj_pos=[[1.5, 1.0, -1.0]])
    g.close()
    assert not np.array_equal(moved, ref)
    S = orc.Scene(example_scene)
    S.set_frame(example_scene["camera"], example_scene["lights"], obj_pos=[[1.5, 1.0, -1.0]])
    rgb_o, _, _, _ = S.render_region(64, 48)
    assert (np.abs(moved.astype(int) - rgb_o.astype(int)).max(axis=1) <= 1).mean() >= 0.99


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_many_lights_and_objects_use_the_device_frame_path(rt, precision):
    """More than 8 lights / 4 objects do not fit the by-value frame parameters: the library then keeps them in device memory
    (synchronised uploads).  Same parity bar."""
    sc = random_soup_scene(n_tris=120, seed=41, n_obj=3, n_lights=4, with_normals=True)
    rng = np.random.default_rng(3)
    sc["lights"] = np.concatenate([sc["lights"], np.column_stack([rng.uniform(-8, 8, (7, 2)), np.full(7, 9.0), rng.uniform(0, 1, (7, 3)).round(2)])])
    sc["obj_pos"] = np.concatenate([sc["obj_pos"], [[1.6, 1.1, -0.5], [-1.7, -1.2, 0.3], [0.1, 1.9, -2.0]]])
    sc["obj_mesh"] = np.zeros(6, dtype=np.uint32)
    assert sc["lights"].shape[0] == 11 and sc["obj_pos"].shape[0] == 6
    if precision == "exact":
        st = assert_exact(rt, sc, 160, 120)
        assert st["shadow_rays"] == 11 * st["hit_pixels"]
    else:
        within, same, _, _ = fast_fractions(rt, sc, 160, 120)
        assert within >= 0.99 and same >= 0.99
    # moving the objects per frame through the device path
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    moved = sc["obj_pos"] + [0.2, -0.1, 0.05]
    rgb = g.bulk_trace(sc["camera"], sc["lights"], 96, 72, 0, 0, 96, 72, obj_pos=moved)
    g.close()
    S = orc.Scene(sc)
    S.set_frame(sc["camera"], sc["lights"], obj_pos=moved)
    rgb_o, _, _, _ = S.render_region(96, 72)
    assert np.array_equal(rgb, rgb_o)


def test_concurrent_bulk_trace_on_one_handle(rt, example_scene):
    """gRPC serves every BulkTrace on its own goroutine (final_report.pdf p.6): concurrent calls on ONE handle, with different
    frame states and regions, must each get their own consistent result (rt_bulk_trace = frame_set + render under one lock)."""
    from concurrent.futures import ThreadPoolExecutor
    from distributed_raytracer_b200 import workloads
    W, H = 192, 108
    cams = [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(20.0 * k), 0.2 * k, 1.04719755) for k in range(4)]
    rects = [(0, 0, 96, 108), (96, 0, 96, 108), (0, 0, 192, 54), (48, 27, 100, 60)]
    S = orc.Scene(example_scene)
    want = []
    for cam, (x, y, w, h) in zip(cams, rects):
        S.set_frame(cam, example_scene["lights"])
        want.append(S.render_region(W, H, x, y, w, h)[0])
    g = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)

    def call(k):
        out = []
        for _ in range(25):
            x, y, w, h = rects[k]
            out.append(g.bulk_trace(cams[k], example_scene["lights"], W, H, x, y, w, h).copy())
        return out
    with ThreadPoolExecutor(4) as pool:
        got = list(pool.map(call, range(4)))
    g.close()
    for k in range(4):
        for res in got[k]:
            assert np.array_equal(res, want[k])


def test_device_api_lanes_run_frames_concurrently(rt, example_scene):
    """Device API on explicit lanes (RT_LANE(k)): four frames with different cameras in flight on four streams of ONE scene;
    every frame must equal its own oracle image (per-lane frame state and scratch)."""
    import torch
    from distributed_raytracer_b200 import workloads
    W, H = 160, 90
    g = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    assert g.lane_count() >= 5
    cams = [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(25.0 * k), 0.1 * k, 1.04719755) for k in range(4)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = [torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda") for _ in range(4)]
    for rep in range(3):
        for k in range(4):
            g.set_frame_state(rt.FrameState(cams[k], example_scene["lights"]), lane=k + 1)
            g.render_region_device(W, H, 0, 0, W, H, outs[k].data_ptr(), streams[k].cuda_stream, rt.LANE(k + 1))
    torch.cuda.synchronize()
    S = orc.Scene(example_scene)
    for k in range(4):
        S.set_frame(cams[k], example_scene["lights"])
        assert np.array_equal(outs[k].cpu().numpy().reshape(-1, 3), S.render_region(W, H)[0])
    with pytest.raises(rt.RtError):
        g.render_region_device(W, H, 0, 0, W, H, outs[0].data_ptr(), None, rt.LANE(200))
    g.close()


@pytest.mark.parametrize("shared", [True, False])
def test_pipelined_renderer_single_gpu_frames_in_flight(rt, example_scene, shared):
    """TileShardedRenderer at one rank: 7 different cameras pipelined over 3 frames in flight (lanes of ONE scene, or 3 scene
    copies); every frame, device-resident and through the pinned-host path, equals its own region render."""
    import torch
    from distributed_raytracer_b200 import parallel, workloads
    W, H, S = 400, 224, 3
    cams = [workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(31.0 * f), 0.04 * f, 1.04719755) for f in range(7)]
    ref = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    want = []
    for c in cams:
        ref.set_frame(c, example_scene["lights"])
        want.append(ref.render_region(W, H))
    ref.close()
    r = parallel.TileShardedRenderer(example_scene, W, H, precision=rt.FP32_FAST, replicas=S, layout_flags=rt.LAYOUT_COLMAJOR,
                                     shared_scene=shared)
    assert len(set(id(g) for g in r.backend.scenes)) == (1 if shared else S)
    for to_host in (False, True):
        for f0 in range(0, len(cams), S):
            outs = [r.render_frame(rt.FrameState(cams[f], example_scene["lights"]), to_host=to_host) for f in range(f0, min(len(cams), f0 + S))]
            r.wait()
            torch.cuda.synchronize()
            for k, o in enumerate(outs):
                assert np.array_equal(o.cpu().numpy().reshape(-1, 3), want[f0 + k]), (to_host, f0 + k)
    for g in set(r.backend.scenes):
        g.close()


@pytest.mark.parametrize("col,count_key,msg", [(0, "verts", "vertex index"), (4, "normals", "normal index"), (6, "materials", "material index")])
def test_scene_create_rejects_out_of_range_indices(rt, example_scene, col, count_key, msg):
    """The indexed mesh is validated where it is gathered -- on the device (k_flatten_faces): one bad index anywhere in the
    face table fails rt_scene_create with RT_ERR_INVALID and names the kind (like a Go slice index panic would abort register())."""
    import copy
    sc = copy.deepcopy({k: example_scene[k] for k in ("meshes", "obj_pos", "obj_mesh")})
    m = sc["meshes"][0]
    faces = np.array(m["faces"], dtype=np.uint32).reshape(-1, 7)
    faces[len(faces) // 2, col] = np.asarray(m[count_key]).reshape(-1, 3 if count_key != "materials" else 10).shape[0]
    m["faces"] = faces
    with pytest.raises(rt.RtError) as ei:
        rt.GpuScene(sc, precision=rt.FP32_FAST)
    assert ei.value.code == rt._lib.RT_ERR_INVALID and msg in str(ei.value)


@pytest.mark.parametrize("name", ["no_objects", "mesh_without_faces", "single_triangle", "degenerate_triangles", "edge_on_triangle"])
def test_degenerate_inputs(rt, name):
    """Edge-case inputs (conftest.degenerate_scenes): no objects, a mesh without faces, a single triangle, zero-area triangles,
    an edge-on triangle -- at 1x1, 3x5 and 64x48 pixels.  FP64_EXACT bit-identical to the oracle, FP32_FAST within tolerance."""
    sc = degenerate_scenes()[name]
    S = orc.Scene(sc)
    ge = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    gf = rt.GpuScene(sc, precision=rt.FP32_FAST)
    for (W, H) in [(1, 1), (3, 5), (64, 48)]:
        rgb_o, col_o, face_o, st = S.render_region(W, H, mode=orc.CULLED)
        rgb, col, face, obj, lit = ge.render_region_debug(W, H)
        assert np.array_equal(face, face_o) and np.array_equal(col.view(np.uint64), col_o.view(np.uint64)) and np.array_equal(rgb, rgb_o)
        if name in ("no_objects", "mesh_without_faces"):
            assert not rgb.any()
        fast = gf.render_region(W, H)
        assert (np.abs(fast.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=1) <= 1).mean() >= 0.99
    ge.close(); gf.close(); S.close()


def test_exact_applies_the_reference_object_level_predicate(rt):
    """tracer.go:32 + box.go:29-69 on the GPU: FP64_EXACT == the oracle's reference-shaped walk on the stress scene in which the
    reference's non-conservative box predicate rejects the object's bounds for a shadow ray that crosses them
    (tests/golden/box_predicate_object_level.npz; CPU twin: test_emul.py)."""
    from conftest import box_predicate_scene
    sc, W, H = box_predicate_scene()
    S = orc.Scene(sc)
    S.build_rtrees()
    rgb_r, col_r, face_r, _ = S.render_region(W, H, mode=orc.RTREE)
    rgb_c, col_c, face_c, _ = S.render_region(W, H, mode=orc.CULLED)
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    rgb, col, face, obj, lit = g.render_region_debug(W, H)
    g.close(); S.close()
    assert np.array_equal(face, face_r) and np.array_equal(col.view(np.uint64), col_r.view(np.uint64)) and np.array_equal(rgb, rgb_r)
    assert not np.array_equal(col_c.view(np.uint64), col_r.view(np.uint64))


def test_host_tracer_bulk_trace_and_draw(rt, example_scene):
    """The reference-facing interface: Tracer.BulkTrace (distributed/main.go:46-91) over partition() rectangles
    (master/main.go:54-91) reassembles into the same screen as draw() (sequential/main.go:15-32)."""
    from distributed_raytracer_b200 import host
    W, H = 320, 240
    env = host.Environment(EXAMPLE_SCENE)
    tr = host.Tracer(env, W, H, precision=rt.FP64_EXACT)
    rgb_o, _, _, _ = oracle_render(example_scene, W, H, layout=orc.L_ROWMAJOR)
    sdl_o, _, _, _ = oracle_render(example_scene, W, H, layout=orc.L_ROWMAJOR, quant=orc.Q_RGBA16_SHR8)
    screen = np.zeros((W, H, 3), dtype=np.uint8)
    rects = host.partition(W, H, 8)
    assert len(rects) == 8
    for (x, y, w, h) in rects:
        res = tr.BulkTrace(x, y, w, h)                           # results[i*h + j]
        screen[x:x + w, y:y + h] = res.reshape(w, h, 3)
    assert np.array_equal(screen.transpose(1, 0, 2), rgb_o.reshape(H, W, 3))
    assert np.array_equal(tr.draw(), sdl_o.reshape(H, W, 3))     # SDL quantiser, Appendix B-6
    assert np.array_equal(tr.draw(rt.QUANT_RGB255), rgb_o.reshape(H, W, 3))
    assert tr.BulkTrace(0, 0, 0, 0).shape == (0, 3)
    tr.close()
    env.close()
