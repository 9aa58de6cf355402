"""Generates tests/golden/trace_results_*.bin with the google.protobuf runtime (6.33.6 in this image): the serialised
comms.TraceResults message (reference shared/comms/comms.proto:34-41, rebuilt here as a dynamic descriptor -- no protoc
needed) for fixed RGB inputs.  These vectors pin the oracle's encoder and, through it, the GPU encoder.

    python tests/golden/make_trace_results.py        # rewrites the .bin / .rgb files next to this script
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def trace_results_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fdp = descriptor_pb2.FileDescriptorProto(name="comms.proto", package="comms", syntax="proto3")
    m = fdp.message_type.add(name="TraceResults")
    c = m.nested_type.add(name="Colour")
    for k, name in enumerate("rgb"):
        c.field.add(name=name, number=k + 1, type=F.TYPE_UINT32, label=F.LABEL_OPTIONAL)
    m.field.add(name="results", number=1, type=F.TYPE_MESSAGE, type_name=".comms.TraceResults.Colour", label=F.LABEL_REPEATED)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("comms.TraceResults"))


def protobuf_encode(rgb):
    """What the worker does at distributed/main.go:52-54,79-86 (one Colour per pixel, in order) followed by Marshal."""
    TR = trace_results_class()
    t = TR()
    for r, g, b in np.asarray(rgb, dtype=np.uint8).reshape(-1, 3).tolist():
        t.results.add(r=r, g=g, b=b)
    return t.SerializeToString()


def inputs():
    rng = np.random.default_rng(0x5EED)
    edge = np.array([[0, 0, 0], [1, 2, 3], [127, 127, 127], [128, 128, 128], [255, 255, 255], [0, 200, 0], [255, 0, 127], [0, 0, 128],
                     [127, 128, 0]], dtype=np.uint8)
    rnd = rng.integers(0, 256, size=(3001, 3), dtype=np.uint8)
    sparse = rnd.copy()
    sparse[rng.random(3001) < 0.9] = 0                      # mostly background, like a rendered frame
    return {"edge": edge, "random": rnd, "sparse": sparse}


if __name__ == "__main__":
    for name, rgb in inputs().items():
        rgb.tofile(os.path.join(HERE, "trace_results_%s.rgb" % name))
        with open(os.path.join(HERE, "trace_results_%s.bin" % name), "wb") as f:
            f.write(protobuf_encode(rgb))
        print(name, rgb.shape[0], "pixels ->", os.path.getsize(os.path.join(HERE, "trace_results_%s.bin" % name)), "bytes")
