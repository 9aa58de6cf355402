"""Times the TraceResults protobuf encoder (csrc/wire.cu: k_wire_onepass; RT_B200_WIRE_ONEPASS=0: k_wire_sizes, k_wire_scan, k_wire_encode) on a 4K frame of C3-like
content (example scene rendered at 3840x2160: ~10 % hit pixels), L2 flushed before every launch, CUDA events on the launching
stream; prints one JSON line with the algorithmic bytes and the fraction of the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import distributed_raytracer_b200 as rt
from distributed_raytracer_b200 import host

W, H = 3840, 2160
env = host.Environment(os.path.join(ROOT, "tests", "golden", "example", "scene.json"))
sc = env.scene_arrays()
env.close()
g = rt.GpuScene(sc, precision=rt.FP32_FAST)
n = W * H
d_rgb = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
g.render_region_device(W, H, 0, 0, W, H, d_rgb.data_ptr(), None, 0)
cap = 11 * n
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ms = []
reps = int(os.environ.get("WIRE_REPS", 20))
for it in range(3 + reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    g.encode_trace_results_device(d_rgb.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), st.cuda_stream, 0)
    e1.record(st)
    st.synchronize()
    if it >= 3:
        ms.append(e0.elapsed_time(e1))
m = int(d_len.item())
from oracle import oracle as orc                              # checker only
same = bool(np.array_equal(d_out[:m].cpu().numpy(), orc.encode_trace_results(d_rgb.cpu().numpy().reshape(-1, 3))))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
onepass = os.environ.get("RT_B200_WIRE_ONEPASS", "0") not in ("", "0")
alg = (1 if onepass else 2) * 3 * n + m                        # RGB read once (single pass) or by the size pass and the encode pass, + the message written
t = float(np.median(ms))
print(json.dumps({"kernels": "k_wire_onepass (decoupled look-back)" if onepass else "k_wire_sizes + k_wire_scan + k_wire_encode", "pixels": n, "message_bytes": m, "ms": t, "ms_min": float(min(ms)),
                  "algorithmic_bytes": alg, "achieved_gbs": alg / t / 1e6, "peak_gbs": peak, "frac": alg / t / 1e6 / peak,
                  "identical_to_oracle": same, "hit_fraction": float((d_rgb.view(-1, 3).amax(dim=1) > 0).float().mean().item())}))
