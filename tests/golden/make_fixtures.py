"""Copies the reference's example DATA fixtures (scene.json, suzanne.obj, suzanne.mtl -- Blender's stock
Suzanne mesh, not source code) from /root/reference/example into tests/golden/example so that the GPU
box (which has no /root/reference) renders the same scene and camera as the reference would.
Run once in the build container:  python tests/golden/make_fixtures.py
"""
import os
import shutil

SRC = "/root/reference/example"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for f in ("scene.json", "suzanne.obj", "suzanne.mtl"):
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        print("copied", f)
