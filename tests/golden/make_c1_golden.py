"""Generates tests/golden/c1_320x240.npz: the C1 reference image (example/scene.json, 3 lights, 320x240)
rendered by the CPU oracle in BRUTE mode -- hit-face map, float64 colours (as uint64 bit patterns) and
both 8-bit quantisations.  The Go reference cannot run in this image (no Go toolchain), so this file
pins the ORACLE's output against regressions; it is not reference output.
Run: python tests/golden/make_c1_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import scene_loader as sl  # noqa: E402

if __name__ == "__main__":
    sc = sl.environment_from_file(os.path.join(ROOT, "tests", "golden", "example", "scene.json"))
    S = orc.Scene(sc)
    rgb, col, face, st = S.render_region(320, 240, mode=orc.BRUTE, quant=orc.Q_RGB255)
    sdl, _, _, _ = S.render_region(320, 240, mode=orc.CULLED, quant=orc.Q_RGBA16_SHR8)
    hit = face >= 0
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c1_320x240.npz"),
                        face=face.astype(np.int16), hit_index=np.flatnonzero(hit).astype(np.int32),
                        colour_bits=col[hit].view(np.uint64), rgb=rgb[hit], sdl=sdl[hit],
                        stats=np.array([st["primary_rays"], st["shadow_rays"], st["hit_pixels"]], dtype=np.int64))
    print(st)
