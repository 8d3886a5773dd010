"""Host-side mesh construction restated from UI::init_dsc / set_dsc_boundary_layer / update_flag."""
import numpy as np

import golden_io


def _rows(a):
    a = np.sort(np.asarray(a, np.int64), axis=1)
    return a[np.lexsort(a.T[::-1])]


def test_grid_mesh_matches_reference_init_dsc(pkg):
    g = golden_io.load("c1_e6")
    pos, tets, grid = pkg.synthetic.grid_mesh((100, 100, 100), 6)
    assert grid == (20, 20, 20)
    assert np.array_equal(pos, g["pos"])  # bit-identical node positions
    assert np.array_equal(_rows(tets), _rows(g["tet_nodes"]))
    base = pkg.synthetic.boundary_layer_labels(tets, grid)
    assert int((base != 0).sum()) == 29478  # tets with label != 0


def test_jittered_grid_matches_reference_driver_topology(pkg):
    g = golden_io.load("sph48_f32")
    pos, tets, grid = pkg.synthetic.grid_mesh((48, 48, 48), 4.3)
    assert np.array_equal(_rows(tets), _rows(g["tet_nodes"]))
    assert len(pos) == len(g["pos"])


def test_interface_face_enumeration_matches_reference(pkg):
    for name in golden_io.CASES:
        g = golden_io.load(name)
        fn, ft, fb = pkg.synthetic.interface_faces(g["tet_nodes"], g["tet_label"], len(g["pos"]))
        assert np.array_equal(_rows(fn), _rows(g["face_nodes"])), name
        assert int(fb.sum()) == int(g["face_is_boundary"].sum())
        # same co-boundary tets per face
        key = lambda n: {tuple(sorted(r)) for r in n}
        mine = {tuple(sorted(n)): tuple(sorted(t)) for n, t in zip(fn.tolist(), ft.tolist())}
        ref = {tuple(sorted(n)): tuple(sorted(t)) for n, t in zip(g["face_nodes"].tolist(), g["face_tets"].tolist())}
        assert mine == ref, name


def test_phantoms_are_deterministic(pkg):
    a = pkg.synthetic.shells_phantom(32, sigma=0.05, dtype=np.uint16)
    b = pkg.synthetic.shells_phantom(32, sigma=0.05, dtype=np.uint16)
    assert np.array_equal(a, b) and a.dtype == np.uint16 and len(np.unique(a)) > 100
    s = pkg.synthetic.sphere_phantom(32, 10)
    assert set(np.unique(s).tolist()) == {np.float32(0.2).item(), np.float32(0.8).item()}
