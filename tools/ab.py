"""A/B timing of kernel-variant libraries (build_variants/*/librt_b200.so): one subprocess per variant."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:] or sorted(os.listdir(os.path.join(ROOT, "build_variants")))
for n in names:
    env = dict(os.environ, RT_B200_LIB=os.path.join(ROOT, "build_variants", n, "librt_b200.so"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "explore.py"), "quick", "none"], env=env, capture_output=True, text=True)
    for l in out.stdout.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            print("%-10s prim %.3f shade %.3f write %.3f total %.3f  Mrays/s %.0f  nodes/ray p %.2f s %.1f  checksum %d" % (
                n, d["ms"]["primary"], d["ms"]["shade"], d["ms"]["write"], d["ms"]["total"], d["mrays_s"], d["nodes_primary_per_ray"],
                d["nodes_shadow_per_ray"], d["checksum"]), flush=True)
    if out.returncode != 0:
        print(n, "FAILED", out.stderr[-500:])
