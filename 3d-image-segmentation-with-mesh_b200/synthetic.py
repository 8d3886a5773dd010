"""Synthetic phantoms and flat mesh construction for tests and the benchmark (host side, numpy).

grid_mesh restates the reference's initial mesh generator UI::init_dsc (DEMO/user_interface.cpp:
550-630: regular grid, 6 tets per cube, padded by one edge length) and boundary_layer_labels its
UI::set_dsc_boundary_layer (:441-462); interface_faces enumerates interface faces from the flat
tet list with the predicate of ISMesh::update_flag(FaceKey) (is_mesh/is_mesh.h:453-477).  Face
and tet ORDER of a flat mesh is this module's own convention (the reference's order is the key
order of its is_mesh complex, which the C++ adapter captures directly); both the CUDA path and
the oracle consume the same arrays, so parity does not depend on it.
"""
import numpy as np

NO_TET = 0xFFFFFFFF
_TETRAS = np.array([0, 4, 5, 7, 0, 7, 5, 1, 0, 1, 3, 7, 1, 5, 6, 7, 1, 6, 7, 3, 1, 2, 6, 3]).reshape(6, 4)


def _cround(x):
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def grid_mesh(dims_xyz, edge, jitter=0.0, seed=1234):
    """Returns pos [n,3] float64, tets [m,4] uint32, (NX,NY,NZ).  dims_xyz = image size (X,Y,Z)."""
    dsc_dim = [d + 2 * edge for d in dims_xyz]
    N = [_cround(d / edge) + 1 for d in dsc_dim]
    delta = [dsc_dim[i] / float(N[i] - 1) for i in range(3)]
    NX, NY, NZ = N
    iz, iy, ix = np.meshgrid(np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    pos = np.stack([ix * delta[0] - edge, iy * delta[1] - edge, iz * delta[2] - edge], axis=-1).reshape(-1, 3).astype(np.float64)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        j = rng.uniform(-1.0, 1.0, pos.shape) * (jitter * edge)
        interior = ((ix >= 2) & (iy >= 2) & (iz >= 2) & (ix <= NX - 3) & (iy <= NY - 3) & (iz <= NZ - 3)).reshape(-1)
        pos[interior] += j[interior]
    cz, cy, cx = np.meshgrid(np.arange(NZ - 1), np.arange(NY - 1), np.arange(NX - 1), indexing="ij")
    cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)

    def idx(x, y, z):
        return z * NX * NY + y * NX + x

    corners = np.stack([idx(cx, cy, cz), idx(cx + 1, cy, cz), idx(cx + 1, cy + 1, cz), idx(cx, cy + 1, cz), idx(cx, cy, cz + 1),
                        idx(cx + 1, cy, cz + 1), idx(cx + 1, cy + 1, cz + 1), idx(cx, cy + 1, cz + 1)], axis=1)
    tets = corners[:, _TETRAS].reshape(-1, 4).astype(np.uint32)
    return pos, tets, (NX, NY, NZ)


def boundary_layer_labels(tets, n_grid):
    """Label 0 for every tet that touches a node of the mesh boundary, 1 otherwise."""
    NX, NY, NZ = n_grid
    n = np.arange(NX * NY * NZ)
    ix, iy, iz = n % NX, (n // NX) % NY, n // (NX * NY)
    outer = (ix == 0) | (iy == 0) | (iz == 0) | (ix == NX - 1) | (iy == NY - 1) | (iz == NZ - 1)
    return np.where(outer[tets].any(axis=1), 0, 1).astype(np.int32)


def interface_faces(tets, labels, n_nodes):
    """Interface faces of a flat tet mesh.

    A face is an interface when its two tets carry different labels, or it has a single tet whose
    label is not 0 (is_mesh/is_mesh.h:453-477).  Returns face_nodes [k,3] (ascending node ids),
    face_tets [k,2] (ascending tet index, NO_TET when single), face_is_boundary [k] -- ordered by
    the face's sorted node triple.
    """
    tets = np.asarray(tets, np.int64)
    m = len(tets)
    loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    tri = np.sort(tets[:, loc].reshape(-1, 3), axis=1)
    owner = np.repeat(np.arange(m, dtype=np.int64), 4)
    N = np.int64(n_nodes)
    if float(N) ** 3 < 9e18:
        key = (tri[:, 0] * N + tri[:, 1]) * N + tri[:, 2]
        order = np.argsort(key, kind="stable")
    else:  # > 2 M nodes: the triple no longer fits one 63-bit key
        order = np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))
    tri, owner = tri[order], owner[order]
    first = np.ones(len(tri), bool)
    first[1:] = (tri[1:] != tri[:-1]).any(axis=1)
    key = tri  # only its length is used below
    start = np.flatnonzero(first)
    count = np.diff(np.append(start, len(key)))
    assert count.max() <= 2, "non-manifold face"
    t0 = owner[start]
    t1 = np.where(count == 2, owner[np.minimum(start + 1, len(key) - 1)], -1)
    lab = np.asarray(labels)
    l0 = lab[t0]
    l1 = np.where(t1 >= 0, lab[np.maximum(t1, 0)], 0)
    interface = np.where(t1 >= 0, l0 != l1, l0 != 0)
    sel = np.flatnonzero(interface)
    face_nodes = tri[start[sel]].astype(np.uint32)
    a, b = t0[sel], t1[sel]
    lo = np.where(b >= 0, np.minimum(a, b), a)
    hi = np.where(b >= 0, np.maximum(a, b), NO_TET)
    face_tets = np.stack([lo, hi], axis=1).astype(np.uint32)
    face_is_boundary = (b < 0).astype(np.uint8)
    return face_nodes, face_tets, face_is_boundary


def _radius(n, centre=None):
    c = (n - 1) / 2.0 if centre is None else centre
    z, y, x = np.ogrid[0:n, 0:n, 0:n]
    return np.sqrt(((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2).astype(np.float32))


def sphere_phantom(n=128, radius=40.0, lo=0.2, hi=0.8):
    """BASELINE config 2: float32 sphere, two phases, no noise."""
    return np.where(_radius(n) < radius, np.float32(hi), np.float32(lo)).astype(np.float32)


def _lobed_radius(n):
    # nested "hamster-like" shells: radius modulated by a smooth angular term so that interfaces are not spherical
    c = (n - 1) / 2.0
    z, y, x = np.ogrid[0:n, 0:n, 0:n]
    dx, dy, dz = ((x - c) / n).astype(np.float32), ((y - c) / n).astype(np.float32), ((z - c) / n).astype(np.float32)
    r = np.sqrt(dx * dx + dy * dy + dz * dz)
    mod = 1.0 + 0.12 * np.sin(9.0 * dx + 1.0) * np.cos(7.0 * dy) + 0.08 * np.sin(11.0 * dz)
    return (r / mod.astype(np.float32)).astype(np.float32)


def shells_phantom(n, levels=(0.1, 0.5, 0.9), radii=(0.40, 0.24), sigma=0.0, seed=1234, dtype=np.float32):
    """Nested shells: levels[0] outside, then one level per radius (fractions of n), optional Gaussian noise
    clipped to [0,1].  dtype float32 keeps the values; uint16 stores round(v*65535).  Built slab by slab so that
    512^3 and 1024^3 volumes do not need float64 temporaries."""
    out = np.empty((n, n, n), dtype=dtype)
    rng = np.random.default_rng(seed)
    slab = max(1, min(n, (1 << 24) // (n * n)))
    c = (n - 1) / 2.0
    y, x = np.ogrid[0:n, 0:n]
    dx, dy = ((x - c) / n).astype(np.float32), ((y - c) / n).astype(np.float32)
    for z0 in range(0, n, slab):
        z1 = min(n, z0 + slab)
        dz = ((np.arange(z0, z1, dtype=np.float32) - c) / n).reshape(-1, 1, 1)
        r = np.sqrt(dx * dx + dy * dy + dz * dz)
        mod = 1.0 + 0.12 * np.sin(9.0 * dx + 1.0) * np.cos(7.0 * dy) + 0.08 * np.sin(11.0 * dz)
        r = r / mod
        v = np.full(r.shape, levels[0], np.float32)
        for lev, rad in zip(levels[1:], radii):
            v = np.where(r < rad, np.float32(lev), v)
        if sigma > 0:
            v = np.clip(v + rng.standard_normal(v.shape, dtype=np.float32) * np.float32(sigma), 0.0, 1.0)
        if np.dtype(dtype) == np.uint16:
            out[z0:z1] = np.round(v * 65535.0).astype(np.uint16)
        elif np.dtype(dtype) == np.uint8:
            out[z0:z1] = np.round(v * 255.0).astype(np.uint8)
        else:
            out[z0:z1] = v.astype(dtype)
    return out


def labels_from_means(tet_mean, base_labels, thresholds):
    """label = 1 + number of thresholds below the tet mean (same rule as DEMO/segment_function.cpp:250-260 with
    fixed thresholds); boundary-layer tets (base label 0) keep 0."""
    lab = 1 + np.searchsorted(np.asarray(thresholds, np.float64), tet_mean, side="left")
    return np.where(np.asarray(base_labels) == 0, 0, lab).astype(np.int32)


CONFIGS = {
    # name: dims, dtype, generator kwargs, grid cells per axis (N-1), phases, thresholds   
    "C2": dict(n=128, dtype=np.float32, cells=32, nb_phase=2, thresholds=(0.5,)),
    "C3": dict(n=256, dtype=np.uint16, cells=51, nb_phase=3, thresholds=(0.3, 0.7)),
    "C4": dict(n=512, dtype=np.float32, cells=69, nb_phase=3, thresholds=(0.3, 0.7)),
    "C5": dict(n=1024, dtype=np.uint16, cells=139, nb_phase=4, thresholds=(0.25, 0.5, 0.75)),
}


def make_volume(name):
    cfg = CONFIGS[name]
    n = cfg["n"]
    if name == "C2":
        return sphere_phantom(n, radius=40.0)
    if name == "C3":
        return shells_phantom(n, levels=(0.1, 0.5, 0.9), radii=(0.40, 0.24), sigma=0.05, seed=1234, dtype=np.uint16)
    if name == "C4":
        return shells_phantom(n, levels=(0.1, 0.5, 0.9), radii=(0.40, 0.24), sigma=0.02, seed=1234, dtype=np.float32)
    if name == "C5":
        return shells_phantom(n, levels=(0.1, 0.4, 0.65, 0.9), radii=(0.42, 0.30, 0.16), sigma=0.03, seed=1234, dtype=np.uint16)
    raise KeyError(name)


def edge_for_cells(n, cells):
    """Edge length such that init_dsc produces `cells` cubes per axis: (n + 2e)/e = cells."""
    return n / float(cells - 2)


def build_case(volume, edge, nb_phase, thresholds, tet_mean_fn, jitter=0.0, seed=1234):
    """Mesh + labels + interface faces for a volume.  tet_mean_fn(pos, tets) -> per-tet mean intensity
    (the GPU's tet_integrate in the benchmark, the oracle in CPU tests)."""
    n = volume.shape[0]
    pos, tets, grid = grid_mesh((volume.shape[2], volume.shape[1], volume.shape[0]), edge, jitter, seed)
    base = boundary_layer_labels(tets, grid)
    mean = tet_mean_fn(pos, tets, base)
    labels = labels_from_means(mean, base, thresholds)
    fn, ft, fb = interface_faces(tets, labels, len(pos))
    return {"pos": pos, "tet_nodes": tets, "tet_label": labels, "face_nodes": fn, "face_tets": ft, "face_is_boundary": fb}
