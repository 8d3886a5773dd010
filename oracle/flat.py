"""ctypes front-end of the flat CPU oracle (oracle/flat_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (it fails loudly without its CUDA library).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DTYPE_CODE = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}


def build(force=False):
    so = os.path.join(_HERE, "libflatoracle.so")
    src = os.path.join(_HERE, "flat_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "flat"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _dims(vol):
    # volume arrays are indexed [z, y, x] (x fastest), DEMO/image3d.h:66-68
    return (C.c_int * 3)(vol.shape[2], vol.shape[1], vol.shape[0])


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return lib().fo_num_threads()


def decode(vol):
    """voxel decode rule (DEMO/image3d.cpp:127 for bytes)."""
    if vol.dtype == np.uint8:
        return vol.astype(np.float64) / 255.0
    if vol.dtype == np.uint16:
        return vol.astype(np.float64) / 65535.0
    return vol.astype(np.float64)


def tet_integrate(vol, tet_nodes, pos, c=None, tet_mask=None, threads=1, want_hash=True):
    vol = np.ascontiguousarray(vol)
    tn = _c(tet_nodes, np.uint32)
    pos = _c(pos, np.float64)
    n = tn.shape[0]
    nc = 0 if c is None else len(c)
    cc = None if c is None else _c(c, np.float64)
    out = {
        "total": np.full(n, -1.0), "vol": np.full(n, -1.0), "mean": np.full(n, -1.0),
        "sq": np.full((n, nc), -1.0) if nc else None, "ad": np.full((n, nc), -1.0) if nc else None,
        "hash": np.zeros(n, np.uint64) if want_hash else None, "nsamp": np.zeros(n, np.int32),
    }
    m = None if tet_mask is None else _c(tet_mask, np.uint8)
    r = lib().fo_tet_integrate(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), C.c_uint32(n), _p(tn), _p(pos), _p(m), C.c_int(nc), _p(cc),
                               _p(out["total"]), _p(out["vol"]), _p(out["mean"]), _p(out["sq"]), _p(out["ad"]), _p(out["hash"]),
                               _p(out["nsamp"]), C.c_int(threads))
    assert r == 0
    return out


def phase_reduce(tet_label, tet_total, tet_vol, nb_phase):
    lab = _c(tet_label, np.int32)
    tt = _c(tet_total, np.float64)
    tv = _c(tet_vol, np.float64)
    total = np.zeros(nb_phase)
    volume = np.zeros(nb_phase)
    mean = np.zeros(nb_phase)
    lib().fo_phase_reduce(C.c_uint32(len(lab)), _p(lab), _p(tt), _p(tv), C.c_int(nb_phase), _p(total), _p(volume), _p(mean))
    return total, volume, mean


def face_forces(vol, snap, nb_phase, mean, threads=1):
    vol = np.ascontiguousarray(vol)
    pos = _c(snap["pos"], np.float64)
    nnb = pos.size // 3
    forces = np.zeros(3 * nnb)
    fn = _c(snap["face_nodes"], np.uint32)
    ft = _c(snap["face_tets"], np.uint32)
    tn = _c(snap["tet_nodes"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    mean = _c(mean, np.float64)
    lib().fo_face_forces(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), C.c_uint32(nnb), _p(pos), _p(tn), _p(tl), C.c_uint32(fn.size // 3), _p(fn),
                         _p(ft), C.c_int(nb_phase), _p(mean), _p(forces), C.c_int(threads))
    return forces.reshape(-1, 3)


def thicken_forces(vol, snap, nb_phase, mean, vb):
    """Force loop of segment_function::thickenning_surface (DEMO/segment_function.cpp:1990-2051): (forces, face_mean, face_dev, evaluated)."""
    vol = np.ascontiguousarray(vol)
    pos = _c(snap["pos"], np.float64)
    nnb = pos.size // 3
    fn = _c(snap["face_nodes"], np.uint32)
    ft = _c(snap["face_tets"], np.uint32)
    tn = _c(snap["tet_nodes"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    nf = fn.size // 3
    forces = np.zeros(3 * nnb)
    mu, dev, ev = np.zeros(nf), np.zeros(nf), np.zeros(nf, np.uint8)
    lib().fo_thicken_forces(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), C.c_uint32(nnb), _p(pos), _p(tn), _p(tl), C.c_uint32(nf), _p(fn), _p(ft),
                            _p(_c(vb, np.uint8)), C.c_int(nb_phase), _p(_c(mean, np.float64)), _p(forces), _p(mu), _p(dev), _p(ev))
    return forces.reshape(-1, 3), mu, dev, ev


def init_histogram(tet_label, tet_mean):
    """Histogram of initialization_discrete_opt (DEMO/segment_function.cpp:219-226) over the labelled tets."""
    m = np.asarray(tet_label) != 0
    idx = np.minimum(np.trunc(np.asarray(tet_mean, np.float64)[m] * 256).astype(np.int64), 255)
    return np.bincount(idx, minlength=256).astype(np.int32)


def init_labels(tet_label, tet_mean, thresholds):
    """Label rule of initialization_discrete_opt (DEMO/segment_function.cpp:245-261): 0 stays 0."""
    m255 = np.minimum(np.trunc(np.asarray(tet_mean, np.float64) * 255).astype(np.int64), 255)
    lab = 1 + np.searchsorted(np.asarray(thresholds, np.int64), m255, side="left")
    return np.where(np.asarray(tet_label) != 0, lab, 0).astype(np.int32)


def internal_forces(snap, node_flags, vb=None, order=None):
    """segment_function::compute_internal_force_2 (DEMO/segment_function.cpp:2197-2232) over a flat snapshot.
    order = (node_face_off, node_face_idx): the interface faces of get_faces(node) per node (bit-identical to the reference then)."""
    pos = _c(snap["pos"], np.float64)
    nnb = pos.size // 3
    fn = _c(snap["face_nodes"], np.uint32)
    ft = _c(snap["face_tets"], np.uint32)
    tn = _c(snap["tet_nodes"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    vb = vertex_bound(snap) if vb is None else _c(vb, np.uint8)
    nf = _c(node_flags, np.uint8)
    out = np.zeros(3 * nnb)
    off = idx = None
    if order is not None:
        off, idx = _c(order[0], np.uint32), _c(order[1], np.uint32)
        assert off.size == nnb + 1 and int(off[-1]) == idx.size
    lib().fo_internal_forces(C.c_uint32(nnb), _p(pos), _p(tn), _p(tl), C.c_uint32(fn.size // 3), _p(fn), _p(ft), _p(vb), _p(nf),
                             _p(off), _p(idx), _p(out))
    return out.reshape(-1, 3)


def build_sum_table(vol):
    vol = np.ascontiguousarray(vol)
    table = np.empty(vol.shape, np.float64)
    lib().fo_build_sum_table(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), _p(table))
    return table


def vertex_bound(snap):
    pos = _c(snap["pos"], np.float64)
    nnb = pos.size // 3
    fn = _c(snap["face_nodes"], np.uint32)
    ft = _c(snap["face_tets"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    vb = np.zeros(nnb, np.uint8)
    lib().fo_vertex_bound(C.c_uint32(nnb), _p(tl), C.c_uint32(fn.size // 3), _p(fn), _p(ft), _p(vb))
    return vb


def zray(vol, snap, nb_phase, mode=0, sum_table=None, mean=None, alpha=0.0, vb=None):
    vol = np.ascontiguousarray(vol)
    pos = _c(snap["pos"], np.float64)
    fn = _c(snap["face_nodes"], np.uint32)
    ft = _c(snap["face_tets"], np.uint32)
    fb = _c(snap["face_is_boundary"], np.uint8)
    tn = _c(snap["tet_nodes"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    if mode == 0 and sum_table is None:
        sum_table = build_sum_table(vol)
    if mode == 1 and vb is None:
        vb = vertex_bound(snap)
    total = np.zeros(nb_phase)
    count = np.zeros(nb_phase)
    mean_out = np.zeros(nb_phase)
    e_data = C.c_double(0)
    e_area = C.c_double(0)
    n_rays = C.c_int64(0)
    hist = np.zeros(64, np.int32)
    mi = None if mean is None else _c(mean, np.float64)
    lib().fo_zray(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), _p(sum_table), _p(pos), _p(tn), _p(tl), C.c_uint32(fn.size // 3), _p(fn), _p(ft),
                  _p(fb), C.c_int(nb_phase), C.c_int(mode), _p(mi), C.c_double(alpha), _p(vb), _p(total), _p(count), _p(mean_out),
                  C.byref(e_data), C.byref(e_area), C.byref(n_rays), _p(hist))
    return {"total": total, "count": count, "mean": mean_out, "energy_data": e_data.value, "energy_area": e_area.value,
            "n_rays": n_rays.value, "hits_hist": hist}


def points_in_tets(tet_idx, pts, tet_nodes, pos):
    ti = _c(tet_idx, np.uint32)
    pts = _c(pts, np.float64)
    tn = _c(tet_nodes, np.uint32)
    pos = _c(pos, np.float64)
    out = np.zeros(len(ti), np.uint8)
    lib().fo_points_in_tets(C.c_uint32(len(ti)), _p(ti), _p(pts), _p(tn), _p(pos), _p(out))
    return out


def std_sort_hits(z, b_in):
    z = _c(z, np.int32).copy()
    b = _c(b_in, np.uint8).copy()
    lib().fo_std_sort_hits(C.c_int(len(z)), _p(z), _p(b))
    return z, b


def relabel_energies(vol, snap, tet_face_nodes, tet_face_tets, nb_phase, mean, alpha):
    vol = np.ascontiguousarray(vol)
    pos = _c(snap["pos"], np.float64)
    tn = _c(snap["tet_nodes"], np.uint32)
    tl = _c(snap["tet_label"], np.int32)
    fn = _c(tet_face_nodes, np.uint32)
    ft = _c(tet_face_tets, np.uint32)
    mean = _c(mean, np.float64)
    out = np.empty((len(tl), nb_phase))
    lib().fo_relabel_energies(_dims(vol), DTYPE_CODE[vol.dtype], _p(vol), C.c_uint32(len(tl)), _p(tn), _p(tl), _p(pos), _p(fn), _p(ft),
                              C.c_int(nb_phase), _p(mean), C.c_double(alpha), _p(out))
    return out
