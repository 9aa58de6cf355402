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


# ------------------------------------------------------------------------------------------------ master-side decoder (CPU)
def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_decoder_library_exports_declared_symbols(built):
    import ctypes as C
    import re
    from conftest import ROOT
    from distributed_raytracer_b200 import wire
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rt_wire.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(rt_wire_[a-z_0-9]+)\s*\(", src)))
    assert set(names) == set(wire.WIRE_SYMBOLS)
    for path in (wire.LIB_PATH, os.path.join(os.path.dirname(wire.LIB_PATH), "librt_host.so")):
        lib = C.CDLL(path)
        for n in names:
            assert getattr(lib, n) is not None


def test_decoder_inverts_protobuf_and_oracle_encoders(built, orc):
    """decode(protobuf.SerializeToString(m)) == m.results and decode(oracle encoder(x)) == x, incl. every byte value."""
    from distributed_raytracer_b200 import wire
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 300, 4097):
        rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
        rgb[rng.random(n) < 0.5] = 0
        assert np.array_equal(wire.decode_trace_results(orc.encode_trace_results(rgb)), rgb)
        try:
            enc = mtr.protobuf_encode(rgb)
        except ImportError:
            continue
        assert np.array_equal(wire.decode_trace_results(enc), rgb)
    for name in ("edge", "random", "sparse"):
        rgb = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.rgb" % name), dtype=np.uint8).reshape(-1, 3)
        enc = np.fromfile(os.path.join(GOLDEN, "trace_results_%s.bin" % name), dtype=np.uint8)
        assert np.array_equal(wire.decode_trace_results(enc), rgb)


def test_decoder_accepts_any_valid_encoding_like_protobuf_does():
    """Non-canonical but valid encodings: fields out of order, repeated scalars (last wins), unknown fields of every wire
    type at both levels, non-minimal varints, values above 255 (the master truncates: uint8(pixel.GetR()), main.go:172).
    Each is also parsed with the protobuf runtime to make sure the expectation is protobuf's, not ours."""
    from distributed_raytracer_b200 import wire
    sub1 = b"\x18\x07" + b"\x08\x05" + b"\x10\x06"                           # b, r, g out of order
    sub2 = b"\x08\x01" + b"\x08\x09"                                          # r twice: last wins
    sub3 = b"\x08" + varint(300) + b"\x10" + varint(2 ** 32 - 1)                # 300 -> 44, 0xffffffff -> 255
    sub4 = b"\x08\x83\x80\x00" + b"\x20\x05" + b"\x2a\x02hi" + b"\x35\x01\x02\x03\x04" + b"\x39" + bytes(8) + b"\x18\x04"
    msg = b"".join(b"\x0a" + varint(len(s)) + s for s in (sub1, sub2, sub3, sub4))
    sub5 = b"\x08\x02" + b"\x43" + b"\x08\x63" + b"\x44" + b"\x1d\x01\x02\x03\x04" + b"\x10\x03"   # an unknown group (proto2 leftover) is
    msg = b"".join(b"\x0a" + varint(len(s)) + s for s in (sub1, sub2, sub3, sub4, sub5))              # skipped; b as fixed32 = unknown
    msg = b"\x10\x2a" + msg + b"\x1a\x03abc" + b"\x0a\x00"                     # unknown top-level fields + an empty Colour
    want = np.array([[5, 6, 7], [9, 0, 0], [44, 255, 0], [3, 0, 4], [2, 3, 0], [0, 0, 0]], dtype=np.uint8)
    assert np.array_equal(wire.decode_trace_results(msg), want)
    try:
        m = mtr.trace_results_class()()
    except ImportError:
        return
    m.ParseFromString(msg)
    assert np.array_equal(np.array([[c.r & 255, c.g & 255, c.b & 255] for c in m.results], dtype=np.uint8), want)


def test_decoder_rejects_malformed_input_and_respects_capacity():
    from distributed_raytracer_b200 import wire
    good = b"\x0a\x02\x08\x01" * 3
    assert wire.decode_trace_results(good).shape == (3, 3)
    for bad in (good[:-1], b"\x0a\x05\x08\x01", b"\x0a", b"\x0b\x00", b"\x08" + b"\xff" * 11, b"\x0a\x02\x08\x80", b"\x00\x00",
                b"\x88\x80\x80\x80\x10\x00",              # a tag wider than 32 bits
                b"\x0a\x02\x43\x4c",                        # group closed by another field's end-group
                b"\x0c"):                                    # stray end-group
        with pytest.raises(wire.WireError) as ei:
            wire.decode_trace_results(bad)
        assert ei.value.code == wire.ERR_MALFORMED, bad
    with pytest.raises(wire.WireError) as ei:
        wire.decode_trace_results(good, max_results=2)
    assert ei.value.code == wire.ERR_CAPACITY


def test_blit_is_the_masters_frame_assembly(orc, example_oracle):
    """rt_wire_blit_trace_results == the loop of master/main.go:166-176: two orders (partition rectangles) of one frame
    assembled into a 32-bit surface; equals the frame rendered in one piece, row-major."""
    from distributed_raytracer_b200 import wire
    W, H = 96, 64
    whole = example_oracle.render_region(W, H, layout=orc.L_ROWMAJOR)[0].reshape(H, W, 3).astype(np.uint32)
    want = (whole[..., 0] << 16) | (whole[..., 1] << 8) | whole[..., 2] | 0xFF000000
    surface = np.zeros((H, W + 5), dtype=np.uint32)[:, :W]                  # a pitch larger than the width
    for (x, y, w, h) in [(0, 0, 40, H), (40, 0, W - 40, H)]:
        rgb = example_oracle.render_region(W, H, x, y, w, h)[0]             # BulkTrace order i*h + j
        wire.blit_trace_results(orc.encode_trace_results(rgb), x, y, w, h, surface, amask=0xFF000000)
    assert np.array_equal(surface, want)
    with pytest.raises(wire.WireError):                                     # a reply with too few results for its order
        wire.blit_trace_results(orc.encode_trace_results(np.zeros((10, 3), np.uint8)), 0, 0, 4, 4, surface)
    with pytest.raises(wire.WireError):
        wire.blit_trace_results(b"", 90, 0, 10, 10, surface)                # region outside the surface


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
