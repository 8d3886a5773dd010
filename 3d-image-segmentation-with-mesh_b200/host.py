"""ctypes host side of the C ABI (include/dscgpu.h), mirroring the reference's call surface.

The reference reaches this path through direct member calls on one ``segment_function`` object
that owns an ``image3d`` (DEMO/segment_function.h:87-218, DEMO/image3d.h:29-86).  The two
classes below keep those names, argument meanings and result fields so the parity tests read
like calls into the reference:

    img = Image3D(volume)                   # image3d::load / _voxels / build_sum_table
    seg = SegmentFunction(img, nb_phase=3)  # segment_function with NB_PHASE, m_alpha
    seg.set_snapshot(snap)                  # the flattened is_mesh/DSC state (see dscgpu_snapshot)
    seg.update_average_intensity()          # -> _mean_intensities, _total_intensities, _phase_volume
    seg.compute_external_force()            # -> _forces   [n_nodes_buffer, 3]
    seg.compute_energy()                    # -> (data term, alpha * area term)

Everything computes on the GPU through libdscgpu.so; a missing library or device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_PHASE = 8
NO_TET = 0xFFFFFFFF
FORCE_GATHER, FORCE_ATOMIC = 0, 1
MEAN_TET, MEAN_ZRAY = 0, 1
DTYPE_CODE = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
TIMERS = {"tet": 0, "phase_final": 1, "force": 2, "zray": 3, "sumtable": 4, "h2d": 5, "d2h": 6}


class DscGpuError(RuntimeError):
    pass


class _Snapshot(C.Structure):
    _fields_ = [("n_nodes_buffer", C.c_uint32), ("pos", C.c_void_p), ("n_tets", C.c_uint32), ("tet_nodes", C.c_void_p),
                ("tet_label", C.c_void_p), ("n_faces", C.c_uint32), ("face_nodes", C.c_void_p), ("face_tets", C.c_void_p),
                ("face_is_boundary", C.c_void_p)]


class _Staging(C.Structure):
    _fields_ = [("pos4", C.c_void_p), ("tet_nodes", C.c_void_p), ("tet_label", C.c_void_p), ("face_nodes", C.c_void_p),
                ("face_tets", C.c_void_p), ("face_is_boundary", C.c_void_p)]


class _UpdateStaging(C.Structure):
    _fields_ = [("node_keys", C.c_void_p), ("pos", C.c_void_p), ("tet_index", C.c_void_p), ("tet_nodes", C.c_void_p), ("tet_label", C.c_void_p),
                ("face_nodes", C.c_void_p), ("face_tets", C.c_void_p), ("face_is_boundary", C.c_void_p)]


class _PhaseStats(C.Structure):
    _fields_ = [("total", C.c_double * MAX_PHASE), ("sumsq", C.c_double * MAX_PHASE), ("volume", C.c_double * MAX_PHASE),
                ("mean", C.c_double * MAX_PHASE)]


class _MoveParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("alpha", C.c_double), ("align_threshold", C.c_double), ("max_displacement", C.c_double)]


class _MoveStats(C.Structure):
    _fields_ = [("max_dis", C.c_double), ("max_boundary_dis", C.c_double), ("scale", C.c_double), ("real_max", C.c_double)]


def library_path():
    return os.path.join(_HERE, "libdscgpu.so")


def load_library():
    """Loads libdscgpu.so, building it with nvcc first if it is missing or stale.  Never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _build
    path = library_path()
    try:
        # development: DSCGPU_LIBRARY names an already built experimental copy of the library (kernel A/B runs)
        path = os.environ["DSCGPU_LIBRARY"] if os.environ.get("DSCGPU_LIBRARY") else _build.build()
    except Exception as e:  # no nvcc on this machine: use a prebuilt library if it is there
        if not os.path.exists(path):
            raise DscGpuError("libdscgpu.so is missing and cannot be built (%s); there is no CPU fallback" % e)
    lib = C.CDLL(path)
    lib.dscgpu_last_error.restype = C.c_char_p
    lib.dscgpu_launch_count.restype = C.c_uint64
    lib.dscgpu_last_sample_count.restype = C.c_uint64
    lib.dscgpu_last_gauss_count.restype = C.c_uint64
    lib.dscgpu_volume_ptr_dev.restype = C.c_void_p
    lib.dscgpu_sum_table_ptr_dev.restype = C.c_void_p
    lib.dscgpu_forces_pinned.restype = C.POINTER(C.c_double)
    lib.dscgpu_comm_window_dev.restype = C.c_void_p
    lib.dscgpu_comm_window_dev.argtypes = [C.c_void_p]
    for name in ("dscgpu_ctx_destroy", "dscgpu_last_error", "dscgpu_set_stream", "dscgpu_synchronize", "dscgpu_build_sum_table",
                 "dscgpu_commit_snapshot", "dscgpu_launch_count", "dscgpu_last_sample_count", "dscgpu_last_gauss_count",
                 "dscgpu_volume_ptr_dev", "dscgpu_sum_table_ptr_dev"):
        getattr(lib, name).argtypes = [C.c_void_p] + ([C.c_void_p] if name == "dscgpu_set_stream" else [])
    _LIB = lib
    return lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def host_sort_hits(z, b_in):
    """Host instance of the device-side restatement of libstdc++ std::sort used by the z-ray kernel."""
    z = np.ascontiguousarray(z, np.int32).copy()
    b = np.ascontiguousarray(b_in, np.uint8).copy()
    r = load_library().dscgpu_host_sort_hits(C.c_int(len(z)), _ptr(z), _ptr(b))
    if r != 0:
        raise DscGpuError("dscgpu_host_sort_hits failed")
    return z, b


class Image3D:
    """Device-resident voxel volume: the GPU counterpart of ``image3d`` (DEMO/image3d.h:29-86).

    ``volume`` is indexed [z, y, x] (x fastest, DEMO/image3d.h:66-68) with dtype uint8, uint16,
    float32 or float64; see include/dscgpu.h for the decode rule.
    """

    def __init__(self, volume, device=0):
        volume = np.ascontiguousarray(volume)
        if volume.ndim != 3 or volume.dtype not in DTYPE_CODE:
            raise ValueError("volume must be a 3-D array of uint8/uint16/float32/float64")
        self.lib = load_library()
        self.shape = volume.shape
        self.dtype = volume.dtype
        self._dim = (volume.shape[2], volume.shape[1], volume.shape[0])
        self.device = device
        self._ctx = C.c_void_p()
        dims = (C.c_int * 3)(*self._dim)
        r = self.lib.dscgpu_ctx_create(C.byref(self._ctx), C.c_int(device), dims, C.c_int(DTYPE_CODE[volume.dtype]), _ptr(volume))
        if r != 0:
            msg = self.lib.dscgpu_last_error(self._ctx).decode() if self._ctx else "context allocation failed"
            if self._ctx:
                self.lib.dscgpu_ctx_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise DscGpuError("dscgpu_ctx_create: " + msg)
        self._has_table = False

    def _check(self, r, what):
        if r != 0:
            raise DscGpuError("%s failed (%d): %s" % (what, r, self.lib.dscgpu_last_error(self._ctx).decode()))

    def dimension(self):
        return self._dim

    def build_sum_table(self):
        """image3d::build_sum_table (DEMO/image3d.cpp:143-177)."""
        self._check(self.lib.dscgpu_build_sum_table(self._ctx), "dscgpu_build_sum_table")
        self._has_table = True

    def sum_table(self):
        out = np.empty(self.shape, np.float64)
        self._check(self.lib.dscgpu_get_sum_table(self._ctx, _ptr(out)), "dscgpu_get_sum_table")
        return out

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.dscgpu_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr)), "dscgpu_set_stream")

    def synchronize(self):
        self._check(self.lib.dscgpu_synchronize(self._ctx), "dscgpu_synchronize")

    def volume_ptr_dev(self):
        return self.lib.dscgpu_volume_ptr_dev(self._ctx)

    def write_volume(self, volume):
        """Replaces the voxels (same shape and dtype): image3d::load on a live object."""
        volume = np.ascontiguousarray(volume)
        if volume.shape != self.shape or volume.dtype != self.dtype:
            raise ValueError("volume must keep its shape and dtype")
        self._check(self.lib.dscgpu_volume_upload(self._ctx, _ptr(volume)), "dscgpu_volume_upload")
        self._has_table = False

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.dscgpu_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SegmentFunction:
    """GPU counterpart of the image-energy members of ``segment_function`` (DEMO/segment_function.h:87-218)."""

    def __init__(self, img, nb_phase, alpha=0.1):
        if not 1 <= nb_phase <= MAX_PHASE:
            raise ValueError("NB_PHASE must be in 1..%d" % MAX_PHASE)
        self._img = img
        self.lib = img.lib
        self.NB_PHASE = nb_phase
        self.m_alpha = alpha
        self._mean_intensities = None
        self._total_intensities = None
        self._phase_volume = None
        self._phase_sumsq = None
        self._forces = None
        self._internal_forces = None
        self.n_nodes_buffer = self.n_tets = self.n_faces = 0
        self._keep = None

    # -- snapshot ---------------------------------------------------------------------------
    def set_snapshot(self, snap):
        """Uploads the flattened mesh (keys: pos, tet_nodes, tet_label, face_nodes, face_tets, face_is_boundary)."""
        pos = np.ascontiguousarray(snap["pos"], np.float64).reshape(-1, 3)
        tn = np.ascontiguousarray(snap["tet_nodes"], np.uint32).reshape(-1, 4)
        tl = np.ascontiguousarray(snap["tet_label"], np.int32).reshape(-1)
        fn = np.ascontiguousarray(snap["face_nodes"], np.uint32).reshape(-1, 3)
        ft = np.ascontiguousarray(snap["face_tets"], np.uint32).reshape(-1, 2)
        fb = np.ascontiguousarray(snap["face_is_boundary"], np.uint8).reshape(-1)
        if len(tl) != len(tn) or len(ft) != len(fn) or len(fb) != len(fn):
            raise ValueError("inconsistent snapshot array lengths")
        s = _Snapshot(len(pos), pos.ctypes.data, len(tn), tn.ctypes.data, tl.ctypes.data, len(fn), fn.ctypes.data, ft.ctypes.data,
                      fb.ctypes.data)
        self._img._check(self.lib.dscgpu_set_snapshot(self._img._ctx, C.byref(s)), "dscgpu_set_snapshot")
        self.n_nodes_buffer, self.n_tets, self.n_faces = len(pos), len(tn), len(fn)

    def update_snapshot(self, node_keys=None, pos=None, tet_index=None, tet_nodes=None, tet_label=None, faces=None):
        """Edits the snapshot resident on the device: moved nodes (keys + positions; keys None = nodes 0..n-1), tets replaced in
        place (index, 4 node keys, label), and -- faces = dict(face_nodes, face_tets, face_is_boundary) -- a new interface-face list."""
        n_pos = 0 if pos is None else len(pos)
        p = None if pos is None else np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        k = None if node_keys is None else np.ascontiguousarray(node_keys, np.uint32)
        if k is not None and len(k) != n_pos:
            raise ValueError("node_keys and pos differ in length")
        n_t = 0 if tet_index is None else len(tet_index)
        ti = None if tet_index is None else np.ascontiguousarray(tet_index, np.uint32)
        tn = None if tet_nodes is None else np.ascontiguousarray(tet_nodes, np.uint32).reshape(-1, 4)
        tl = None if tet_label is None else np.ascontiguousarray(tet_label, np.int32)
        if n_t and (tn is None or tl is None or len(tn) != n_t or len(tl) != n_t):
            raise ValueError("tet_index, tet_nodes and tet_label differ in length")
        fn = ft = fb = None
        nf = 0
        if faces is not None:
            fn = np.ascontiguousarray(faces["face_nodes"], np.uint32).reshape(-1, 3)
            ft = np.ascontiguousarray(faces["face_tets"], np.uint32).reshape(-1, 2)
            fb = np.ascontiguousarray(faces["face_is_boundary"], np.uint8).reshape(-1)
            nf = len(fn)
        r = self.lib.dscgpu_update_snapshot(self._img._ctx, C.c_uint32(n_pos), _ptr(k), _ptr(p), C.c_uint32(n_t), _ptr(ti), _ptr(tn), _ptr(tl),
                                            C.c_int(0 if faces is None else 1), C.c_uint32(nf), _ptr(fn) if nf else None, _ptr(ft) if nf else None,
                                            _ptr(fb) if nf else None)
        self._img._check(r, "dscgpu_update_snapshot")
        if faces is not None:
            self.n_faces = nf

    def update_staging(self, n_pos, with_keys, n_tet_upd, n_faces):
        """Pinned host arrays of an incremental update (zero-copy); fill them, then update_commit()."""
        st = _UpdateStaging()
        self._img._check(self.lib.dscgpu_update_staging_get(self._img._ctx, C.c_uint32(n_pos), C.c_int(1 if with_keys else 0), C.c_uint32(n_tet_upd),
                                                            C.c_uint32(n_faces), C.byref(st)), "dscgpu_update_staging_get")

        def view(ptr, ctype, n, shape):
            if n == 0 or not ptr:
                return np.zeros(shape, dtype=ctype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape)

        self._upd_faces = n_faces
        return {"node_keys": view(st.node_keys, C.c_uint32, n_pos if with_keys else 0, (n_pos if with_keys else 0,)),
                "pos": view(st.pos, C.c_double, 3 * n_pos, (n_pos, 3)),
                "tet_index": view(st.tet_index, C.c_uint32, n_tet_upd, (n_tet_upd,)),
                "tet_nodes": view(st.tet_nodes, C.c_uint32, 4 * n_tet_upd, (n_tet_upd, 4)),
                "tet_label": view(st.tet_label, C.c_int32, n_tet_upd, (n_tet_upd,)),
                "face_nodes": view(st.face_nodes, C.c_uint32, 3 * n_faces, (n_faces, 3)),
                "face_tets": view(st.face_tets, C.c_uint32, 2 * n_faces, (n_faces, 2)),
                "face_is_boundary": view(st.face_is_boundary, C.c_uint8, n_faces, (n_faces,))}

    def update_commit(self, replace_faces=True):
        self._img._check(self.lib.dscgpu_update_commit(self._img._ctx, C.c_int(1 if replace_faces else 0)), "dscgpu_update_commit")
        if replace_faces:
            self.n_faces = self._upd_faces

    def staging(self, n_nodes_buffer, n_tets, n_faces):
        """Pinned host arrays to flatten a mesh into without an extra copy; follow with commit_snapshot()."""
        st = _Staging()
        self._img._check(self.lib.dscgpu_snapshot_staging_get(self._img._ctx, C.c_uint32(n_nodes_buffer), C.c_uint32(n_tets),
                                                              C.c_uint32(n_faces), C.byref(st)), "dscgpu_snapshot_staging_get")

        def view(ptr, ctype, n, shape):
            if n == 0:
                return np.zeros(shape, dtype=ctype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape)

        self.n_nodes_buffer, self.n_tets, self.n_faces = n_nodes_buffer, n_tets, n_faces
        return {"pos4": view(st.pos4, C.c_double, 4 * n_nodes_buffer, (n_nodes_buffer, 4)),
                "tet_nodes": view(st.tet_nodes, C.c_uint32, 4 * n_tets, (n_tets, 4)),
                "tet_label": view(st.tet_label, C.c_int32, n_tets, (n_tets,)),
                "face_nodes": view(st.face_nodes, C.c_uint32, 3 * n_faces, (n_faces, 3)),
                "face_tets": view(st.face_tets, C.c_uint32, 2 * n_faces, (n_faces, 2)),
                "face_is_boundary": view(st.face_is_boundary, C.c_uint8, n_faces, (n_faces,))}

    def commit_snapshot(self):
        self._img._check(self.lib.dscgpu_commit_snapshot(self._img._ctx), "dscgpu_commit_snapshot")

    def set_partition(self, tet_range, face_range, node_range):
        self._img._check(self.lib.dscgpu_set_partition(self._img._ctx, *[C.c_uint32(int(v)) for v in (*tet_range, *face_range, *node_range)]),
                         "dscgpu_set_partition")

    # -- per-tet integration ------------------------------------------------------------------
    def _stats_out(self, st):
        P = self.NB_PHASE
        self._total_intensities = np.array(st.total[:P])
        self._phase_sumsq = np.array(st.sumsq[:P])
        self._phase_volume = np.array(st.volume[:P])
        self._mean_intensities = np.array(st.mean[:P])

    def tet_integrate(self, c=None, per_tet=True, include_bound=False, update_means=True):
        """image3d::get_tetra_intensity (+ get_energy/get_variation for every c[k]) over all tets and the
        per-label reduction.  Returns a dict of per-tet arrays and the per-phase statistics."""
        nt = self.n_tets
        nc = 0 if c is None else len(c)
        cc = None if c is None else np.ascontiguousarray(c, np.float64)
        out = {}
        if per_tet:
            out["total"] = np.empty(nt)
            out["vol"] = np.empty(nt)
            out["mean"] = np.empty(nt)
            if nc:
                out["energy"] = np.empty((nt, nc))
                out["variation"] = np.empty((nt, nc))
        st = _PhaseStats()
        r = self.lib.dscgpu_tet_integrate(self._img._ctx, C.c_int(self.NB_PHASE), C.c_int(1 if include_bound else 0), C.c_int(nc), _ptr(cc),
                                          _ptr(out.get("total")), _ptr(out.get("vol")), _ptr(out.get("mean")), _ptr(out.get("energy")),
                                          _ptr(out.get("variation")), C.byref(st))
        self._img._check(r, "dscgpu_tet_integrate")
        P = self.NB_PHASE
        out["phase_total"] = np.array(st.total[:P])
        out["phase_sumsq"] = np.array(st.sumsq[:P])
        out["phase_volume"] = np.array(st.volume[:P])
        out["phase_mean"] = np.array(st.mean[:P])
        if update_means:
            self._stats_out(st)
        return out

    def relabel_energies(self, tet_face_nodes, tet_face_tets, mean=None):
        """segment_function::get_energy_tetrahedron (DEMO/segment_function.cpp:618-643) for every labelled tet and every
        candidate label: [n_tets, NB_PHASE]; -1 for label-0 tets."""
        fn = np.ascontiguousarray(tet_face_nodes, np.uint32).reshape(-1)
        ft = np.ascontiguousarray(tet_face_tets, np.uint32).reshape(-1)
        if fn.size != 12 * self.n_tets or ft.size != 8 * self.n_tets:
            raise ValueError("tet_face_nodes / tet_face_tets must describe 4 faces per tet")
        self._img._check(self.lib.dscgpu_set_tet_faces(self._img._ctx, _ptr(fn), _ptr(ft)), "dscgpu_set_tet_faces")
        m = np.ascontiguousarray(self._mean_intensities if mean is None else mean, np.float64)
        out = np.empty((self.n_tets, self.NB_PHASE))
        self._img._check(self.lib.dscgpu_relabel_energies(self._img._ctx, C.c_int(self.NB_PHASE), _ptr(m), C.c_double(self.m_alpha), _ptr(out)),
                         "dscgpu_relabel_energies")
        return out

    def tet_sample_hash(self):
        h = np.zeros(self.n_tets, np.uint64)
        n = np.zeros(self.n_tets, np.int32)
        self._img._check(self.lib.dscgpu_tet_sample_hash(self._img._ctx, _ptr(h), _ptr(n)), "dscgpu_tet_sample_hash")
        return h, n

    def points_in_tets(self, tet_idx, pts):
        ti = np.ascontiguousarray(tet_idx, np.uint32)
        p = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.zeros(len(ti), np.uint8)
        self._img._check(self.lib.dscgpu_points_in_tets(self._img._ctx, C.c_uint32(len(ti)), _ptr(ti), _ptr(p), _ptr(out)), "dscgpu_points_in_tets")
        return out

    # -- the reference's members --------------------------------------------------------------
    def update_average_intensity(self):
        """segment_function::update_average_intensity (DEMO/segment_function.cpp:1600-1764), z-ray estimator."""
        if not self._img._has_table:
            self._img.build_sum_table()
        st = _PhaseStats()
        self._img._check(self.lib.dscgpu_zray_means(self._img._ctx, C.c_int(self.NB_PHASE), C.byref(st)), "dscgpu_zray_means")
        self._stats_out(st)
        return self._mean_intensities

    def compute_external_force(self, mode=FORCE_GATHER, mean=None):
        """segment_function::compute_external_force (DEMO/segment_function.cpp:1425-1473)."""
        m = np.ascontiguousarray(self._mean_intensities if mean is None else mean, np.float64)
        if len(m) != self.NB_PHASE:
            raise ValueError("need NB_PHASE mean intensities")
        f = np.empty((self.n_nodes_buffer, 3))
        self._img._check(self.lib.dscgpu_face_forces(self._img._ctx, C.c_int(self.NB_PHASE), _ptr(m), C.c_int(mode), _ptr(f)), "dscgpu_face_forces")
        self._forces = f
        return f

    def init_histogram(self):
        """Histogram of the per-tet means for Otsu's thresholding (DEMO/segment_function.cpp:165-226)."""
        h = np.zeros(256, np.int32)
        self._img._check(self.lib.dscgpu_init_histogram(self._img._ctx, C.c_int(self.NB_PHASE), _ptr(h)), "dscgpu_init_histogram")
        return h

    def init_labels(self, thresholds):
        """Labels of initialization_discrete_opt for the given Otsu thresholds (DEMO/segment_function.cpp:245-261)."""
        t = np.ascontiguousarray(thresholds, np.int32)
        out = np.zeros(self.n_tets, np.int32)
        self._img._check(self.lib.dscgpu_init_labels(self._img._ctx, C.c_int(len(t)), _ptr(t) if len(t) else None, _ptr(out)), "dscgpu_init_labels")
        return out

    def set_node_face_order(self, node_off, face_idx):
        """The order in which DSC::get_normal(node) (src/DSC.h:3138-3159) adds the face normals: per node key the interface faces of
        get_faces(node) (is_mesh/is_mesh.h:691-699) as indices into the snapshot's face list, CSR.  None drops it."""
        if node_off is None:
            self._img._check(self.lib.dscgpu_set_node_face_order(self._img._ctx, C.c_uint32(0), None, None), "dscgpu_set_node_face_order")
            return
        off = np.ascontiguousarray(node_off, np.uint32)
        idx = np.ascontiguousarray(face_idx, np.uint32)
        if len(off) != self.n_nodes_buffer + 1:
            raise ValueError("node_off must have n_nodes_buffer + 1 entries")
        self._img._check(self.lib.dscgpu_set_node_face_order(self._img._ctx, C.c_uint32(len(idx)), _ptr(off), _ptr(idx) if len(idx) else None),
                         "dscgpu_set_node_face_order")

    def compute_internal_force_2(self, node_flags, vertex_bound=None, face_order=None):
        """segment_function::compute_internal_force_2 (DEMO/segment_function.cpp:2197-2232).  node_flags: bit 0 is_interface,
        bit 1 is_crossing per node (is_mesh attributes).  face_order = (node_off, face_idx): see set_node_face_order; with it every
        force is bit-identical to the reference's, without it the projected ones agree to rounding."""
        nf = np.ascontiguousarray(node_flags, np.uint8)
        if len(nf) != self.n_nodes_buffer:
            raise ValueError("node_flags must have one byte per node of the buffer")
        if face_order is not None:
            self.set_node_face_order(*face_order)
        vb = None if vertex_bound is None else np.ascontiguousarray(vertex_bound, np.uint8)
        f = np.empty((self.n_nodes_buffer, 3))
        self._img._check(self.lib.dscgpu_internal_forces(self._img._ctx, _ptr(nf), _ptr(vb), _ptr(f)), "dscgpu_internal_forces")
        self._internal_forces = f
        return f

    def get_node_displacements(self, dt, forces=None, internal_forces=None):
        """segment_function::get_node_displacement for every node (DEMO/segment_function.cpp:581-584)."""
        F = self._forces if forces is None else forces
        Fi = self._internal_forces if internal_forces is None else internal_forces
        return (F + self.m_alpha * Fi) * dt

    def thicken_forces(self, mean=None, vertex_bound=None):
        """Force loop of segment_function::thickenning_surface (DEMO/segment_function.cpp:1990-2051):
        returns (forces[n,3], face_mean[n_faces], face_dev[n_faces], face_evaluated[n_faces])."""
        m = np.ascontiguousarray(self._mean_intensities if mean is None else mean, np.float64)
        vb = None if vertex_bound is None else np.ascontiguousarray(vertex_bound, np.uint8)
        F = np.zeros((self.n_nodes_buffer, 3))
        mu, dev, ev = np.zeros(self.n_faces), np.zeros(self.n_faces), np.zeros(self.n_faces, np.uint8)
        r = self.lib.dscgpu_thicken_forces(self._img._ctx, C.c_int(self.NB_PHASE), _ptr(m), _ptr(vb), _ptr(F), _ptr(mu), _ptr(dev), _ptr(ev))
        self._img._check(r, "dscgpu_thicken_forces")
        return F, mu, dev, ev

    def node_destinations(self, dt, align_threshold, max_displacement, forces=None, internal_forces=None, node_flags=None, vertex_bound=None, dt_adapt=None):
        """get_node_displacement + the non-topological part of snapp_boundary (DEMO/segment_function.cpp:466-541, 581-584) on the device.
        None for forces / internal_forces / node_flags / vertex_bound = what the last force calls left resident.
        Returns (destinations[n,3], cur_dis[n,3], stats dict)."""
        n = self.n_nodes_buffer
        mp = _MoveParams(float(dt), float(self.m_alpha), float(align_threshold), float(max_displacement))
        st = _MoveStats()
        dest = np.empty((n, 3))
        cur = np.empty((n, 3))
        arr = lambda a, dt_: None if a is None else np.ascontiguousarray(a, dt_)
        F, Fi, nf, vb, da = arr(forces, np.float64), arr(internal_forces, np.float64), arr(node_flags, np.uint8), arr(vertex_bound, np.uint8), arr(dt_adapt, np.float64)
        self._keep_move = (F, Fi, nf, vb, da)
        r = self.lib.dscgpu_node_destinations(self._img._ctx, C.byref(mp), _ptr(F), _ptr(Fi), _ptr(nf), _ptr(vb), _ptr(da), _ptr(dest), _ptr(cur), C.byref(st))
        self._img._check(r, "dscgpu_node_destinations")
        return dest, cur, {"max_dis": st.max_dis, "max_boundary_dis": st.max_boundary_dis, "scale": st.scale, "real_max": st.real_max}

    def compute_energy(self, mean=None, vertex_bound=None):
        """segment_function::compute_energy (DEMO/segment_function.cpp:2331-2492): (data, alpha*area)."""
        m = np.ascontiguousarray(self._mean_intensities if mean is None else mean, np.float64)
        vb = None if vertex_bound is None else np.ascontiguousarray(vertex_bound, np.uint8)
        d = C.c_double(0)
        a = C.c_double(0)
        self._img._check(self.lib.dscgpu_zray_energy(self._img._ctx, C.c_int(self.NB_PHASE), _ptr(m), C.c_double(self.m_alpha), _ptr(vb),
                                                     C.byref(d), C.byref(a)), "dscgpu_zray_energy")
        return d.value, a.value

    def vertex_bound(self):
        vb = np.zeros(self.n_nodes_buffer, np.uint8)
        self._img._check(self.lib.dscgpu_vertex_bound(self._img._ctx, _ptr(vb)), "dscgpu_vertex_bound")
        return vb

    def image_force_eval(self, snap=None, mean_mode=MEAN_TET, force_mode=FORCE_GATHER, want_forces=True):
        """One full evaluation through host buffers: upload + means + forces + download."""
        sp = None
        if snap is not None:
            pos = np.ascontiguousarray(snap["pos"], np.float64).reshape(-1, 3)
            tn = np.ascontiguousarray(snap["tet_nodes"], np.uint32).reshape(-1, 4)
            tl = np.ascontiguousarray(snap["tet_label"], np.int32).reshape(-1)
            fn = np.ascontiguousarray(snap["face_nodes"], np.uint32).reshape(-1, 3)
            ft = np.ascontiguousarray(snap["face_tets"], np.uint32).reshape(-1, 2)
            fb = np.ascontiguousarray(snap["face_is_boundary"], np.uint8).reshape(-1)
            self._keep = (pos, tn, tl, fn, ft, fb)
            sp = _Snapshot(len(pos), pos.ctypes.data, len(tn), tn.ctypes.data, tl.ctypes.data, len(fn), fn.ctypes.data, ft.ctypes.data,
                           fb.ctypes.data)
            self.n_nodes_buffer, self.n_tets, self.n_faces = len(pos), len(tn), len(fn)
        if mean_mode == MEAN_ZRAY and not self._img._has_table:
            self._img.build_sum_table()
        # want_forces: True -> copied into a fresh array; "pinned" -> zero-copy view of the context's pinned result buffer;
        # "compact" -> (node keys, rows) of the nodes with interface faces only, views of the pinned buffers
        self._img._check(self.lib.dscgpu_set_compact_forces(self._img._ctx, C.c_int(1 if want_forces == "compact" else 0)), "dscgpu_set_compact_forces")
        f = np.empty((self.n_nodes_buffer, 3)) if want_forces is True else None
        st = _PhaseStats()
        r = self.lib.dscgpu_image_force_eval(self._img._ctx, C.byref(sp) if sp is not None else None, C.c_int(self.NB_PHASE), C.c_int(mean_mode),
                                             C.c_int(force_mode), C.byref(st), _ptr(f))
        self._img._check(r, "dscgpu_image_force_eval")
        self._stats_out(st)
        if want_forces == "compact" and self.n_nodes_buffer:
            n = C.c_uint32(0)
            keys, rows = C.POINTER(C.c_uint32)(), C.POINTER(C.c_double)()
            self._img._check(self.lib.dscgpu_forces_compact_pinned(self._img._ctx, C.byref(n), C.byref(keys), C.byref(rows)), "dscgpu_forces_compact_pinned")
            k = np.ctypeslib.as_array(keys, shape=(n.value,)) if n.value else np.zeros(0, np.uint32)
            f = (k, np.ctypeslib.as_array(rows, shape=(n.value * 3,)).reshape(-1, 3) if n.value else np.zeros((0, 3)))
            self._forces = None
            return self._mean_intensities, f
        if want_forces == "pinned" and self.n_nodes_buffer:
            n = C.c_uint32(0)
            ptr = self.lib.dscgpu_forces_pinned(self._img._ctx, C.byref(n))
            f = np.ctypeslib.as_array(ptr, shape=(n.value * 3,)).reshape(-1, 3)
        self._forces = f
        return self._mean_intensities, f

    # -- device-resident pieces (multi-GPU composition) ---------------------------------------
    def tet_phase_sums_dev(self, d_sums_ptr):
        self._img._check(self.lib.dscgpu_tet_phase_sums_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_sums_ptr)), "dscgpu_tet_phase_sums_dev")

    def zray_phase_sums_dev(self, d_sums_ptr):
        if not self._img._has_table:
            self._img.build_sum_table()
        self._img._check(self.lib.dscgpu_zray_phase_sums_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_sums_ptr)), "dscgpu_zray_phase_sums_dev")

    def means_from_sums_dev(self, d_sums_ptr, d_mean_ptr):
        self._img._check(self.lib.dscgpu_means_from_sums_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_sums_ptr), C.c_void_p(d_mean_ptr)),
                         "dscgpu_means_from_sums_dev")

    def set_overlap_geometry(self, on=True):
        """Tet passes of the per-iteration entry points start the geometry half of the force pass themselves (dscgpu_set_overlap_geometry)."""
        self._img._check(self.lib.dscgpu_set_overlap_geometry(self._img._ctx, C.c_int(1 if on else 0)), "dscgpu_set_overlap_geometry")

    def face_geometry_async(self):
        """Starts the mean-independent half of the external force on the context's second stream (joined by the next gather-mode force call)."""
        self._img._check(self.lib.dscgpu_face_geometry_async(self._img._ctx), "dscgpu_face_geometry_async")

    def read_forces_compact(self, d_forces_ptr):
        """(node keys, rows) of the nodes with interface faces from a device-resident force array: views of the pinned result buffers."""
        n = C.c_uint32(0)
        keys, rows = C.POINTER(C.c_uint32)(), C.POINTER(C.c_double)()
        self._img._check(self.lib.dscgpu_read_forces_compact(self._img._ctx, C.c_void_p(d_forces_ptr), C.byref(n), C.byref(keys), C.byref(rows)),
                         "dscgpu_read_forces_compact")
        if not n.value:
            return np.zeros(0, np.uint32), np.zeros((0, 3))
        return np.ctypeslib.as_array(keys, shape=(n.value,)), np.ctypeslib.as_array(rows, shape=(n.value * 3,)).reshape(-1, 3)

    def face_forces_dev(self, d_mean_ptr, d_forces_ptr, mode=FORCE_GATHER):
        self._img._check(self.lib.dscgpu_face_forces_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_mean_ptr), C.c_int(mode),
                                                         C.c_void_p(d_forces_ptr)), "dscgpu_face_forces_dev")

    # -- exchange over NVLink peer memory (csrc/dsc_comm.cuh) -------------------------------------------------------
    def comm_create(self, rank, world, cap_rows=None, snapshot_bytes=0):
        """Allocates this rank's exchange window (snapshots of up to cap_rows nodes; snapshot_bytes > 0 adds the region the sharded
        snapshot commit assembles in) and returns its CUDA IPC handle (64 bytes)."""
        cap = int(cap_rows if cap_rows is not None else max(self.n_nodes_buffer, 1))
        h = (C.c_ubyte * 64)()
        self._img._check(self.lib.dscgpu_comm_create_ex(self._img._ctx, C.c_int(rank), C.c_int(world), C.c_uint32(cap), C.c_uint64(int(snapshot_bytes)), h),
                         "dscgpu_comm_create_ex")
        return bytes(h)

    def commit_snapshot_sharded(self):
        """Collective form of commit_snapshot: every rank uploads 1/world of the staged bytes, the rest arrives over NVLink."""
        self._img._check(self.lib.dscgpu_commit_snapshot_sharded(self._img._ctx), "dscgpu_commit_snapshot_sharded")

    def comm_connect(self, handles):
        """handles: the 64-byte IPC handles of all ranks, in rank order (exchanged by the caller)."""
        buf = b"".join(handles)
        self._img._check(self.lib.dscgpu_comm_connect(self._img._ctx, C.c_char_p(buf)), "dscgpu_comm_connect")

    def comm_connect_local(self, windows):
        """Contexts of several ranks inside one process (tests): windows[k] = comm_window() of rank k."""
        arr = (C.c_void_p * len(windows))(*windows)
        self._img._check(self.lib.dscgpu_comm_connect_local(self._img._ctx, arr), "dscgpu_comm_connect_local")

    def comm_window(self):
        return self.lib.dscgpu_comm_window_dev(self._img._ctx)

    def comm_error(self):
        self._img._check(self.lib.dscgpu_comm_error(self._img._ctx), "dscgpu_comm_error")

    def tet_phase_sums_allreduce_dev(self, d_sums_ptr, d_mean_ptr=None):
        self._img._check(self.lib.dscgpu_tet_phase_sums_allreduce_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_sums_ptr),
                                                                      C.c_void_p(d_mean_ptr) if d_mean_ptr else None), "dscgpu_tet_phase_sums_allreduce_dev")

    def face_forces_allreduce_dev(self, d_mean_ptr, d_forces_ptr):
        self._img._check(self.lib.dscgpu_face_forces_allreduce_dev(self._img._ctx, C.c_int(self.NB_PHASE), C.c_void_p(d_mean_ptr),
                                                                   C.c_void_p(d_forces_ptr)), "dscgpu_face_forces_allreduce_dev")

    # -- instrumentation --------------------------------------------------------------------------
    def enable_timers(self, on=True):
        self._img._check(self.lib.dscgpu_enable_timers(self._img._ctx, C.c_int(1 if on else 0)), "dscgpu_enable_timers")

    def timer_ms(self, name):
        ms = C.c_float(0)
        self._img._check(self.lib.dscgpu_get_timer(self._img._ctx, C.c_int(TIMERS[name]), C.byref(ms)), "dscgpu_get_timer")
        return ms.value

    def launch_count(self):
        return int(self.lib.dscgpu_launch_count(self._img._ctx))

    def last_sample_count(self):
        return int(self.lib.dscgpu_last_sample_count(self._img._ctx))

    def last_gauss_count(self):
        return int(self.lib.dscgpu_last_gauss_count(self._img._ctx))

    def set_tet_variant(self, variant):
        """0 v1, 128 w2 (+4096 general widening, +16384 clamp-free batches), 131072 staged (see include/dscgpu.h)."""
        self._img._check(self.lib.dscgpu_set_tet_variant(self._img._ctx, C.c_int(variant)), "dscgpu_set_tet_variant")

    def get_tet_variant(self):
        return int(self.lib.dscgpu_get_tet_variant(self._img._ctx))

    def set_hot_want_sq(self, on):
        """Per-label second moment on the per-iteration entry points (off by default)."""
        self._img._check(self.lib.dscgpu_set_hot_want_sq(self._img._ctx, C.c_int(1 if on else 0)), "dscgpu_set_hot_want_sq")

    def set_lanes_per_tet(self, lanes):
        self._img._check(self.lib.dscgpu_set_lanes_per_tet(self._img._ctx, C.c_int(lanes)), "dscgpu_set_lanes_per_tet")
