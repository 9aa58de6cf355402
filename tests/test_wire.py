"""Result wire format: comms.TraceResults as protobuf bytes.

CPU: the oracle's encoder is PINNED against the google.protobuf runtime (live, and through the committed vectors it produced).
GPU: rt_encode_trace_results_device / rt_bulk_trace_encoded are bit-identical to the oracle (hence to protobuf) on the golden
inputs, on ragged sizes, on a full 4K frame, and the message decodes back to the rendered pixels with the protobuf runtime."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_trace_results as mtr  # noqa: E402


@pytest.fixture(scope="module")
def orc(built):
    from oracle import oracle
    return oracle


@pytest.mark.parametrize("name", ["edge", "random", "sparse"])
def test_oracle_encoder_equals_committed_protobuf_vectors(orc, name):
    rgb = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.rgb" % name), dtype=np.uint8).reshape(-1, 3)
    want = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.bin" % name), dtype=np.uint8)
    assert np.array_equal(rgb, mtr.inputs()[name])                     # the generator still describes the fixture
    assert np.array_equal(orc.encode_trace_results(rgb), want)


def test_oracle_encoder_equals_live_protobuf_runtime(orc):
    pytest.importorskip("google.protobuf")
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 255, 1024, 1025, 4097):
        rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
        rgb[rng.random(n) < 0.5] = 0
        assert orc.encode_trace_results(rgb).tobytes() == mtr.protobuf_encode(rgb), n
    every = np.array([[v, 255 - v, (v * 7) & 255] for v in range(256)], dtype=np.uint8)     # every byte value in every field
    assert orc.encode_trace_results(every).tobytes() == mtr.protobuf_encode(every)


def test_oracle_encoder_decodes_back_with_protobuf(orc, example_oracle):
    """The message built from a rendered region parses, with the protobuf runtime, into the same colours in the same order."""
    pytest.importorskip("google.protobuf")
    rgb = example_oracle.render_region(64, 48)[0]
    msg = mtr.trace_results_class()()
    msg.ParseFromString(orc.encode_trace_results(rgb).tobytes())
    back = np.array([[c.r, c.g, c.b] for c in msg.results], dtype=np.uint8)
    assert np.array_equal(back, rgb.reshape(-1, 3))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def rt(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import distributed_raytracer_b200 as m
    return m


def gpu_encode(rt, g, rgb, capacity=None, lane=0):
    import torch
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    n = rgb.shape[0]
    d_rgb = torch.from_numpy(rgb.reshape(-1).copy()).cuda() if n else torch.zeros(4, dtype=torch.uint8, device="cuda")
    cap = 11 * n if capacity is None else capacity
    d_out = torch.full((max(cap, 1) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
    g.encode_trace_results_device(d_rgb.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), None, rt.LANE(lane))
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), int(d_len.item()), cap


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["edge", "random", "sparse"])
def test_gpu_encoder_equals_protobuf_vectors(rt, example_scene, name):
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    rgb = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.rgb" % name), dtype=np.uint8).reshape(-1, 3)
    want = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.bin" % name), dtype=np.uint8)
    out, n, cap = gpu_encode(rt, g, rgb)
    assert n == want.size and np.array_equal(out[:n], want)
    assert (out[cap:] == 0xEE).all()                                   # nothing written past the buffer
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 3, 1023, 1024, 1025, 2048 + 17, 1024 * 1024 + 5])
def test_gpu_encoder_ragged_sizes_equal_oracle(rt, orc, example_scene, n):
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    rng = np.random.default_rng(n + 1)
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    rgb[rng.random(n) < 0.7] = 0
    want = orc.encode_trace_results(rgb)
    out, m, cap = gpu_encode(rt, g, rgb, lane=1)
    assert m == want.size and np.array_equal(out[:m], want)
    assert (out[max(cap, 1):] == 0xEE).all()
    g.close()


@pytest.mark.gpu
def test_gpu_encoder_reports_length_when_capacity_is_short(rt, orc, example_scene):
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    rgb = np.full((5000, 3), 200, dtype=np.uint8)                      # 11 bytes per pixel
    out, m, cap = gpu_encode(rt, g, rgb, capacity=30000)
    assert m == 55000 and (out[cap:] == 0xEE).all()                    # needed length reported, no byte beyond capacity
    g.close()


@pytest.mark.gpu
def test_bulk_trace_encoded_is_the_protobuf_of_bulk_trace(rt, orc, example_scene):
    """rt_bulk_trace_encoded == protobuf(rt_bulk_trace) for a region of C1, decodes with the protobuf runtime into the pixels;
    a full 4K frame equals the oracle's encoding of the same pixels (size-independent property: decode(encode(x)) == x)."""
    g = rt.GpuScene(example_scene, precision=rt.FP64_EXACT)
    cam, lights = example_scene["camera"], example_scene["lights"]
    W, H = 320, 240
    for (x, y, w, h) in [(0, 0, W, H), (100, 60, 121, 77)]:
        rgb = g.bulk_trace(cam, lights, W, H, x, y, w, h)
        enc = g.bulk_trace_encoded(cam, lights, W, H, x, y, w, h)
        assert np.array_equal(enc, orc.encode_trace_results(rgb))
        try:
            msg = mtr.trace_results_class()()
        except ImportError:
            continue
        msg.ParseFromString(enc.tobytes())
        assert np.array_equal(np.array([[c.r, c.g, c.b] for c in msg.results], dtype=np.uint8), rgb.reshape(-1, 3))
    small = np.empty(100, dtype=np.uint8)
    with pytest.raises(rt.RtError):
        g.bulk_trace_encoded(cam, lights, W, H, 0, 0, W, H, out=small)
    g.close()
    g = rt.GpuScene(example_scene, precision=rt.FP32_FAST)
    W, H = 3840, 2160
    rgb = g.bulk_trace(cam, lights, W, H, 0, 0, W, H)
    enc = g.bulk_trace_encoded(cam, lights, W, H, 0, 0, W, H)
    assert np.array_equal(enc, orc.encode_trace_results(rgb))
    g.close()


@pytest.mark.gpu
def test_host_tracer_bulk_trace_encoded(rt, orc):
    """The C++ mirror of the worker (Tracer.BulkTraceEncoded) returns the protobuf of what Tracer.BulkTrace returns."""
    from conftest import EXAMPLE_SCENE
    from distributed_raytracer_b200 import host
    env = host.Environment(EXAMPLE_SCENE)
    t = host.Tracer(env, 200, 150, precision=rt.FP64_EXACT)
    rgb = t.BulkTrace(30, 20, 101, 67)
    assert np.array_equal(t.BulkTraceEncoded(30, 20, 101, 67), orc.encode_trace_results(rgb))
