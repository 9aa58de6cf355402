"""GPU (-m gpu): BASELINE.json configs 4 and 5 AT SIZE, through the C ABI, against the oracle.

C4: 10 000 000-triangle random soup, 8 white lights on a ring, 3840x2160 (BVH-build and traversal stress; flat-normal path
    object.go:93-95, default material mesh.go:151).  The oracle renders every 256th pixel of the BulkTrace order (32 400 pixels):
    FP64_EXACT must be bit-identical on the sample (face, float64 colour bits, lit mask, bytes), FP32_FAST within 1 LSB on
    >= 99.5 % of it (a soup of sub-pixel triangles is all silhouette; the BASELINE bar of 99.9 % is for the Suzanne configs).
C5: the 120-frame orbit (3 degrees per frame, radius 6) of the subdivided Suzanne, dispatched as master/pool does: the screen is
    cut by partition() (master/main.go:54-91) into one rectangle per worker and every rectangle is one BulkTrace order; four
    angles of the orbit are rendered, assembled like master/main.go:166-176 and compared with the oracle on a pixel sample.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
W, H = 3840, 2160


@pytest.fixture(scope="module")
def rt(built):
    import distributed_raytracer_b200 as rt
    if rt._lib.load().rt_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return rt


def sample_indices(step, phase):
    return np.arange(phase, W * H, step)


def test_c4_ten_million_triangle_soup_at_4k(rt, example_scene):
    from distributed_raytracer_b200 import workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.triangle_soup(10_000_000)]
    sc["lights"] = workloads.ring_lights(8)
    step, phase = 256, 3
    S = orc.Scene(sc)
    rgb_o, col_o, face_o, st = S.render_region(W, H, mode=orc.CULLED, step=step, phase=phase, n_threads=0)
    S.close()
    idx = sample_indices(step, phase)
    assert st["primary_rays"] == len(idx) and st["hit_pixels"] > 10000 and st["shadow_rays"] == 8 * st["hit_pixels"]

    # FP64_EXACT: bit-identical on the sample
    g = rt.GpuScene(sc, precision=rt.FP64_EXACT)
    info = g.bvh_info()
    assert info["n_faces"] == 10_000_000 and 0 < info["n_nodes"] < 10_000_000
    rgb, col, face, obj, lit = g.render_region_debug(W, H)
    g.close()
    assert np.array_equal(face[idx], face_o[idx])
    assert np.array_equal(col[idx].view(np.uint64), col_o[idx].view(np.uint64))
    assert np.array_equal(rgb[idx], rgb_o[idx])
    hits_exact = int((face >= 0).sum())
    del rgb, col, obj, lit

    # FP32_FAST (fused frame kernel, the bench path) and the three-kernel pipeline: same frame; oracle tolerance on the sample
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)
    fast = g.render_region(W, H)
    stats = g.stats()
    pipe = g.render_region(W, H, flags=rt.FLAG_PIPELINE)
    nocull = g.render_region(W, H, flags=rt.FLAG_NO_CULL)
    _, _, face_f, _, _ = g.render_region_debug(W, H)
    g.close()
    assert np.array_equal(fast, pipe) and np.array_equal(fast, nocull)
    within = (np.abs(fast[idx].astype(np.int32) - rgb_o[idx].astype(np.int32)).max(axis=1) <= 1).mean()
    same_face = (face_f[idx] == face_o[idx]).mean()
    assert within >= 0.995 and same_face >= 0.995, (within, same_face)
    assert abs(stats["hit_pixels"] - hits_exact) <= 0.001 * hits_exact
    assert stats["shadow_rays"] == 8 * stats["hit_pixels"]


def test_c5_orbit_through_partition_rectangles(rt, example_scene):
    from distributed_raytracer_b200 import host, workloads
    sc = dict(example_scene)
    sc["meshes"] = [workloads.subdivide(example_scene["meshes"][0], 5)]
    sc["lights"] = example_scene["lights"][:2]
    n_workers = 8
    rects = host.partition(W, H, n_workers)                         # master/main.go:101 -> :54-91
    assert len(rects) == n_workers and sum(w * h for _, _, w, h in rects) == W * H
    g = rt.GpuScene(sc, precision=rt.FP32_FAST)                     # one GPU answers every worker's order here
    S = orc.Scene(sc)
    fov = float(sc["camera"][12])
    step, phase = 512, 7
    idx = sample_indices(step, phase)
    for frame in (0, 17, 60, 101):                                  # 3 degrees per frame
        cam = workloads.orbit_camera([1.0, 1.0, -1.0], 6.0, np.deg2rad(3.0 * frame), 0.0, fov)
        screen = np.zeros((W, H, 3), dtype=np.uint8)                # column-major like the wire layout (i*H + j)
        for (x, y, w, h) in rects:
            out = g.bulk_trace(cam, sc["lights"], W, H, x, y, w, h)   # Tracer.BulkTrace of one WorkOrder
            screen[x:x + w, y:y + h] = out.reshape(w, h, 3)           # master/main.go:166-176
        whole = g.bulk_trace(cam, sc["lights"], W, H, 0, 0, W, H)
        assert np.array_equal(screen.reshape(-1, 3), whole)          # the rectangles tile the frame exactly
        S.set_frame(cam, sc["lights"])
        rgb_o, _, face_o, st = S.render_region(W, H, mode=orc.CULLED, step=step, phase=phase, n_threads=0, want_colour=False)
        assert st["hit_pixels"] > 500
        within = (np.abs(whole[idx].astype(np.int32) - rgb_o[idx].astype(np.int32)).max(axis=1) <= 1).mean()
        assert within >= 0.999, (frame, within)
    S.close()
    g.close()
