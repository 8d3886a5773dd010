"""Reader / writer of the reference's tet-mesh text format (.dsc): is_mesh::import_tet_mesh / export_tet_mesh
(is_mesh/mesh_io.cpp:43-86, 298-314).

    v x y z          one line per vertex, in index order
    t a b c d label  one line per tetrahedron: four vertex indices (get_nodes(t) order) and the label

The reference writes the coordinates with the stream's default precision (6 significant digits), so a file written by it does not
hold the positions it was written from; what a file holds is read back exactly.  The reader is what lets the benchmark replay the
meshes a reference run saves (iter_N.dsc): the interface faces of a loaded mesh are enumerated like those of the synthetic meshes
(synthetic.interface_faces)."""
import numpy as np


def read_dsc(path):
    """Returns (pos[n,3] float64, tets[m,4] uint32, labels[m] int32).  Like the reference's reader, lines are told apart by their
    first character and anything else is ignored."""
    pos, tets, labels = [], [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                pos.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == "t":
                tets.append((int(tok[1]), int(tok[2]), int(tok[3]), int(tok[4])))
                labels.append(int(tok[5]))
    return (np.array(pos, np.float64).reshape(-1, 3), np.array(tets, np.uint32).reshape(-1, 4), np.array(labels, np.int32))


def write_dsc(path, pos, tets, labels, precision=6):
    """Writes what export_tet_mesh writes (precision 6 = the reference's); precision 17 keeps the positions exactly."""
    fmt = "%%.%dg" % precision
    with open(path, "w") as f:
        for p in np.asarray(pos, np.float64).reshape(-1, 3):
            f.write("v %s %s %s\n" % (fmt % p[0], fmt % p[1], fmt % p[2]))
        for t, lab in zip(np.asarray(tets).reshape(-1, 4), np.asarray(labels).reshape(-1)):
            f.write("t %d %d %d %d %d\n" % (t[0], t[1], t[2], t[3], lab))


def snapshot_from_tets(pos, tets, labels):
    """The SoA snapshot of a loaded mesh: interface faces enumerated from the tets and labels (ascending face order, get_nodes(f) /
    get_tets(f) conventions as in synthetic.interface_faces)."""
    from . import synthetic
    fn, ft, fb = synthetic.interface_faces(np.asarray(tets, np.uint32), np.asarray(labels, np.int32), len(pos))
    return {"pos": np.asarray(pos, np.float64), "tet_nodes": np.asarray(tets, np.uint32), "tet_label": np.asarray(labels, np.int32),
            "face_nodes": fn, "face_tets": ft, "face_is_boundary": fb}
