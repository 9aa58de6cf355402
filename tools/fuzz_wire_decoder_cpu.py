"""CPU-only fuzzer: random, deliberately non-canonical and mutated comms.TraceResults messages; rt_wire_decode_trace_results must
accept exactly what the google.protobuf runtime accepts and decode the same values (mod 256, the master's uint8()).
   python tools/fuzz_wire_decoder_cpu.py <seed> <seconds>     Round 1: 2 x 40 s, 215 136 messages, no disagreement (after
the first run taught the decoder 32-bit tags and unknown groups)."""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import numpy as np
import make_trace_results as mtr
from distributed_raytracer_b200 import wire
from google.protobuf.message import DecodeError
TR = mtr.trace_results_class()
rng = np.random.default_rng(int(sys.argv[1]))
def varint(v):
    out = bytearray()
    while v >= 0x80: out.append((v & 0x7f) | 0x80); v >>= 7
    out.append(v); return bytes(out)
def rand_varint_value():
    r = rng.integers(0,6)
    return int([rng.integers(0,128), rng.integers(128,256), rng.integers(256, 70000), 2**32-1, rng.integers(0, 2**63), 0][r])
def rand_field(depth):
    fnum = int(rng.choice([1,2,3,4,5,15,16,100]))
    wt = int(rng.choice([0,0,0,1,2,5]))
    tag = varint((fnum<<3)|wt)
    if wt==0:
        v = varint(rand_varint_value())
        if rng.integers(0,10)==0: v = v[:-1] + bytes([v[-1]|0x80]) + b"\x00"    # non-minimal
        return tag+v
    if wt==1: return tag+bytes(rng.integers(0,256,8,dtype=np.uint8))
    if wt==5: return tag+bytes(rng.integers(0,256,4,dtype=np.uint8))
    body = bytes(rng.integers(0,256,int(rng.integers(0,6)),dtype=np.uint8))
    return tag+varint(len(body))+body
def rand_colour():
    return b"".join(rand_field(1) if rng.integers(0,4)==0 else varint((int(rng.integers(1,4))<<3))+varint(rand_varint_value()) for _ in range(int(rng.integers(0,5))))
t0=time.time(); n=0; bad=0; acc=0
while time.time()-t0 < float(sys.argv[2]):
    parts=[]
    for _ in range(int(rng.integers(0,8))):
        if rng.integers(0,5)==0: parts.append(rand_field(0) if rng.integers(0,2) else b"\x12\x01\x00")
        else:
            c = rand_colour(); parts.append(b"\x0a"+varint(len(c))+c)
    msg = b"".join(parts)
    if rng.integers(0,4)==0 and len(msg)>0:           # mutate: truncate / flip
        k = int(rng.integers(0,len(msg)))
        msg = msg[:k] if rng.integers(0,2) else msg[:k]+bytes([msg[k]^int(rng.integers(1,256))])+msg[k+1:]
    n+=1
    m = TR()
    try:
        m.ParseFromString(msg); pb = np.array([[c.r & 255, c.g & 255, c.b & 255] for c in m.results], dtype=np.uint8).reshape(-1,3); perr=None
    except DecodeError as e: pb=None; perr=e
    try:
        mine = wire.decode_trace_results(msg); merr=None
    except wire.WireError as e: mine=None; merr=e
    if (pb is None) != (mine is None) or (pb is not None and not np.array_equal(pb, mine)):
        bad+=1
        if bad<=8: print("DIFF", msg.hex(), "protobuf:", perr if pb is None else pb.tolist(), "mine:", merr if mine is None else mine.tolist())
    elif pb is not None: acc+=1
print("messages", n, "accepted by both", acc, "bad", bad)
