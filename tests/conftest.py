import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXAMPLE_SCENE = os.path.join(GOLDEN, "example", "scene.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds every native artefact (CUDA library, host mirror, oracle) once per session."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def example_scene():
    from oracle import scene_loader as sl
    return sl.environment_from_file(EXAMPLE_SCENE)


@pytest.fixture(scope="session")
def example_oracle(example_scene):
    from oracle import oracle as orc
    return orc.Scene(example_scene)


def random_soup_scene(n_tris, seed, with_normals=False, n_obj=1, n_lights=2, extent=1.0, half=0.15):
    """Small random scene: float32-origin vertices, integer Ns, several objects sharing one mesh."""
    rng = np.random.default_rng(seed)
    cen = rng.uniform(-extent, extent, size=(n_tris, 1, 3))
    verts = (cen + rng.uniform(-half, half, size=(n_tris, 3, 3))).reshape(-1, 3).astype(np.float32).astype(np.float64)
    faces = np.zeros((n_tris, 7), dtype=np.uint32)
    faces[:, 0:3] = np.arange(3 * n_tris).reshape(-1, 3)
    faces[:, 6] = rng.integers(0, 2, size=n_tris)
    normals = np.zeros((0, 3))
    if with_normals:
        nr = rng.normal(size=(3 * n_tris, 3)).astype(np.float32).astype(np.float64)
        nr /= np.sqrt((nr * nr).sum(axis=1))[:, None]
        normals = nr
        faces[:, 3:6] = np.arange(3 * n_tris).reshape(-1, 3)
    mats = np.array([[0.1, 0.1, 0.1, 0.8, 0.7, 0.6, 0.5, 0.5, 0.5, 10.0],
                     [0.05, 0.1, 0.2, 0.3, 0.9, 0.4, 0.2, 0.3, 0.9, 3.0]]).astype(np.float32).astype(np.float64)
    obj_pos = np.array([[0.0, 0.0, 0.0], [0.7, -0.4, -1.5], [-0.9, 0.5, 0.8]])[:n_obj]
    lights = np.array([[0.0, 0.0, 10.0, 0.0, 1.0, 0.0], [3.0, 10.0, 10.0, 1.0, 0.0, 0.0], [-4.0, -10.0, 6.0, 0.0, 0.0, 1.0],
                       [6.0, 1.0, 8.0, 1.0, 1.0, 1.0]])[:n_lights]
    from oracle import scene_loader as sl
    cam = sl.new_camera([0.2, 0.1, 5.0], [-0.05, -0.02, -1.0], 1.0)
    return {"meshes": [{"verts": verts, "normals": normals, "faces": faces, "materials": mats}],
            "obj_pos": obj_pos, "obj_mesh": np.zeros(n_obj, dtype=np.uint32), "lights": lights, "camera": cam}


def degenerate_scenes():
    """Edge-case inputs (name -> scene): what the reference handles without special code must not trip the GPU path either --
    no objects at all, a mesh without faces (object.go:72 `if m != nil` loop simply finds nothing), a single triangle,
    zero-area and duplicate-vertex triangles (triangle.go:46 incidence == 0 => never hit), a triangle seen exactly edge-on."""
    from oracle import scene_loader as sl
    cam = sl.new_camera([0.0, 0.0, 5.0], [0.0, 0.0, -1.0], 1.0)
    lights = np.array([[0.0, 0.0, 10.0, 1.0, 1.0, 1.0], [5.0, 5.0, 5.0, 0.5, 0.25, 1.0]])
    mat = np.array([[0.1, 0.2, 0.3, 0.8, 0.7, 0.6, 0.5, 0.5, 0.5, 4.0]])

    def mesh(verts, faces):
        f = np.zeros((len(faces), 7), dtype=np.uint32)
        if len(faces):
            f[:, 0:3] = np.asarray(faces, dtype=np.uint32)
        return {"verts": np.asarray(verts, dtype=np.float64).reshape(-1, 3), "normals": np.zeros((0, 3)), "faces": f, "materials": mat}

    def scene(meshes, obj_pos, obj_mesh):
        return {"meshes": meshes, "obj_pos": np.asarray(obj_pos, dtype=np.float64).reshape(-1, 3),
                "obj_mesh": np.asarray(obj_mesh, dtype=np.uint32), "lights": lights, "camera": cam}
    tri = [[-1.0, -1.0, 0.0], [1.0, -1.0, 0.0], [0.0, 1.0, 0.0]]
    # an axis-aligned cube on a 1/8 grid seen by an axis-aligned camera with tan(fov/2) = 1: at 16x16 pixels the pixel centres
    # are odd multiples of 1/16 on the plane one unit ahead, so with the front face 1 unit away rays pass EXACTLY through
    # the shared diagonal, the cube's edges and (for the larger cube) its corners: equal-distance ties between the two
    # triangles of a face (lowest index wins, object.go:97) and zero-thickness boxes in the BVH
    def cube(h):
        v = [[sx * h, sy * h, sz * h] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        q = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
        return mesh(v, [t for a, b, c, d in q for t in ((a, b, c), (c, d, a))])
    cam90 = sl.new_camera([0.0, 0.0, 5.0], [0.0, 0.0, -1.0], 2.0 * np.arctan(1.0))

    def with_cam(sc, c):
        sc = dict(sc)
        sc["camera"] = c
        return sc
    return {
        "cube_edges_and_diagonals": with_cam(scene([cube(0.4375)], [[0.0, 0.0, 3.5625]], [0]), cam90),
        "two_cubes_coplanar_fronts": with_cam(scene([cube(0.3125)], [[-0.3125, 0.0, 3.6875], [0.3125, 0.0, 3.6875]], [0, 0]), cam90),
        "no_objects": scene([], [], []),
        "mesh_without_faces": scene([mesh(tri, [])], [[0, 0, 0]], [0]),
        "single_triangle": scene([mesh(tri, [[0, 1, 2]])], [[0.25, 0.0, 0.0]], [0]),
        "degenerate_triangles": scene([mesh(tri + [[0.5, 0.5, 1.0], [2.0, 2.0, 1.0]],
                                            [[0, 1, 2], [0, 0, 1], [3, 3, 3], [0, 3, 4], [1, 1, 1]])], [[0, 0, 0]], [0]),
        "edge_on_triangle": scene([mesh([[0.0, -1.0, 1.0], [0.0, 1.0, 1.0], [0.0, 0.0, -1.0]] + tri, [[0, 1, 2], [3, 4, 5]])],
                                  [[0, 0, 0], [0.0, 0.0, -2.0]], [0, 0]),
    }


def box_predicate_scene():
    """tests/golden/box_predicate_object_level.npz: a 50-triangle grid-aligned stress scene (found by
    tools/fuzz_degenerate_faces_cpu.py) in which one shadow ray leaves the object's bounds exactly through an edge: box.go's
    predicate (box.go:29-69) rejects the bounds on both adjacent sides by one rounding, so the reference never searches the
    object for that ray (tracer.go:32) and the pixel stays lit although a triangle lies between it and the light."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "box_predicate_object_level.npz"))
    sc = {"meshes": [{k: z[k] for k in ("verts", "normals", "faces", "materials")}], "obj_pos": z["obj_pos"], "obj_mesh": z["obj_mesh"],
          "camera": z["camera"], "lights": z["lights"]}
    return sc, int(z["W"]), int(z["H"])
