"""The .dsc reader (mesh_io.read_dsc, the counterpart of is_mesh::import_tet_mesh, is_mesh/mesh_io.cpp:43-86) against a file written by
the reference's own writer (tests/golden/make_dsc_fixture.py) and the snapshot the same reference run dumped."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _compact(pos, tets):
    """extract_tet_mesh (is_mesh/is_mesh.h:1686-1706) numbers the valid nodes in ascending key order."""
    used = np.zeros(len(pos), bool)
    used[np.unique(tets)] = True
    return used


def test_read_reference_written_file(pkg, tmp_path):
    mio = __import__("dsc_b200.mesh_io", fromlist=["x"])
    pos, tets, lab = mio.read_dsc(os.path.join(HERE, "golden", "deformed_e12.dsc"))
    z = np.load(os.path.join(HERE, "golden", "deformed_e12_dsc.npz"))
    assert tets.shape == z["tet_nodes"].shape and np.array_equal(lab, z["tet_label"])
    # the dump is indexed by raw node key, the file by rank among the valid nodes (garbage_collect + ascending keys)
    valid = np.zeros(len(z["pos"]), bool)
    valid[np.unique(z["tet_nodes"])] = True
    rank = np.cumsum(valid) - 1
    if valid.sum() == len(pos):
        assert np.array_equal(tets, rank[z["tet_nodes"]].astype(np.uint32))
        np.testing.assert_allclose(pos, z["pos"][valid], rtol=1e-5, atol=1e-5)  # the writer keeps 6 significant digits
    else:  # nodes without tets still exist in the complex: the file lists them too
        assert len(pos) >= valid.sum()
    # the format round-trips exactly at full precision, and at the reference's precision the text is reproduced byte for byte
    p17 = tmp_path / "a.dsc"
    mio.write_dsc(p17, pos, tets, lab, precision=17)
    pos2, tets2, lab2 = mio.read_dsc(p17)
    assert np.array_equal(pos2, pos) and np.array_equal(tets2, tets) and np.array_equal(lab2, lab)
    p6 = tmp_path / "b.dsc"
    mio.write_dsc(p6, pos, tets, lab)
    assert open(p6).read() == open(os.path.join(HERE, "golden", "deformed_e12.dsc")).read()


def test_snapshot_of_a_loaded_mesh(pkg):
    mio = __import__("dsc_b200.mesh_io", fromlist=["x"])
    pos, tets, lab = mio.read_dsc(os.path.join(HERE, "golden", "deformed_e12.dsc"))
    snap = mio.snapshot_from_tets(pos, tets, lab)
    nf = len(snap["face_nodes"])
    assert nf > 0 and snap["face_tets"].shape == (nf, 2) and snap["face_is_boundary"].shape == (nf,)
    # every listed face separates two different labels (or is a boundary face of a labelled tet)
    ft = snap["face_tets"]
    two = ft[:, 1] != 0xFFFFFFFF
    assert np.all(lab[ft[two, 0]] != lab[ft[two, 1]])
