/* dscgpu.h -- C ABI of the B200-native DSC image-energy path.
 *
 * Drop-in boundary for the per-iteration image-energy evaluation of the reference
 * (tuannt8/3D-image-segmentation-with-mesh).  The reference has no FFI: the path is reached by
 * direct C++ member calls (DEMO/segment_function.cpp:2242-2247, 2258-2260, 2276-2277).  Each
 * entry point below names the reference member(s) it replaces; the C++ adapter that binds them
 * to a live is_mesh/DSC complex is include/dscgpu_adapter.hpp, the ctypes binding is
 * 3d-image-segmentation-with-mesh_b200/host.py, and INTEGRATION.md shows the patch a maintainer
 * of the reference would apply.
 *
 * Conventions: plain C, opaque context, caller-owned HOST buffers unless the name ends in _dev,
 * int status return (0 = ok, <0 = error; dscgpu_last_error() gives the text), no exceptions
 * cross the boundary, no internal locking (the reference calls this path from one thread only,
 * DEMO/user_interface.cpp:388-396).  There is no CPU fallback: every call needs a CUDA device.
 *
 * Volume layout: x fastest, index = z*X*Y + y*X + x (DEMO/image3d.h:66-68).  Voxel decode:
 *   U8  -> v/255.0 (DEMO/image3d.cpp:127),  U16 -> v/65535.0,  F32 -> widened,  F64 as is;
 * reads outside the volume give 0 (BOUND_INTENSITY, DEMO/image3d.h:24, image3d.cpp:480-491).
 *
 * Labels: tet label 0 = BOUND_LABEL (DEMO/image3d.h:23); phases are labels 1..nb_phase and
 * every per-phase output is indexed by label-1 (DEMO/segment_function.cpp:1445-1446).
 */
#ifndef DSCGPU_H
#define DSCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSCGPU_MAX_PHASE 8
#define DSCGPU_NO_TET 0xFFFFFFFFu

enum dscgpu_dtype { DSCGPU_U8 = 0, DSCGPU_U16 = 1, DSCGPU_F32 = 2, DSCGPU_F64 = 3 };

enum dscgpu_status {
    DSCGPU_OK = 0,
    DSCGPU_ERR_CUDA = -1,        /* a CUDA runtime call failed */
    DSCGPU_ERR_ARG = -2,         /* bad argument */
    DSCGPU_ERR_NO_SNAPSHOT = -3, /* mesh snapshot not set */
    DSCGPU_ERR_NO_SUMTABLE = -4, /* dscgpu_build_sum_table not called */
    DSCGPU_ERR_RAY_OVERFLOW = -5 /* the z-ray hit buffer could not be sized (it is grown and the pass repeated automatically; this is the give-up) */
};

/* rays with more interface hits than this are resolved in place in the hit buffer (slow path) instead of in registers / local memory */
#define DSCGPU_MAX_RAY_HITS 96
/* tet_label of a slot that holds no tet (key-indexed snapshots keep one slot per key of the tet buffer): skipped by every pass, whatever
 * include_bound says; its per-tet outputs are -1 */
#define DSCGPU_LABEL_FREE (-1)

typedef struct dscgpu_ctx dscgpu_ctx;

/* Flattened SoA snapshot of the mesh, taken through the reference's read-only getters in
 * ascending key order (kernel iteration order, is_mesh/kernel.h:178-196):
 *   pos        3*n_nodes_buffer doubles, indexed by raw node key (get_no_nodes_buffer(),
 *              is_mesh/is_mesh.h:150; free cells may hold anything)
 *   tet_nodes  4*n_tets node keys in get_nodes(t) order (is_mesh/is_mesh.h:673-679)
 *   tet_label  n_tets labels (get_label, is_mesh/is_mesh.h:250-253); label-0 tets must be present
 *   face_nodes 3*n_faces node keys in get_nodes(f) order (is_mesh/is_mesh.h:665-671), for every
 *              face with is_interface() (is_mesh/attributes.h:195-252), ascending face key
 *   face_tets  2*n_faces INDICES into the tet arrays, in get_tets(f) order; DSCGPU_NO_TET when
 *              the face has a single co-boundary tet
 *   face_is_boundary  n_faces bytes, FaceAttributes::is_boundary()
 */
typedef struct dscgpu_snapshot {
    uint32_t n_nodes_buffer;
    const double *pos;
    uint32_t n_tets;
    const uint32_t *tet_nodes;
    const int32_t *tet_label;
    uint32_t n_faces;
    const uint32_t *face_nodes;
    const uint32_t *face_tets;
    const uint8_t *face_is_boundary;
} dscgpu_snapshot;

/* Per-phase results.  total/volume/mean are what the reference keeps in _total_intensities,
 * _phase_volume, _mean_intensities (DEMO/segment_function.h:129-131). */
typedef struct dscgpu_phase_stats {
    double total[DSCGPU_MAX_PHASE];
    double sumsq[DSCGPU_MAX_PHASE]; /* sum I^2 * v/n^3 per phase (tet path only) */
    double volume[DSCGPU_MAX_PHASE];
    double mean[DSCGPU_MAX_PHASE];
} dscgpu_phase_stats;

/* -- lifetime ---------------------------------------------------------------------------- */

/* Replaces image3d::load's upload part (DEMO/image3d.cpp:95-133): copies the voxel volume
 * (dims = {X,Y,Z}) to HBM once and installs the sample tables (set_size(),
 * DEMO/tet_dis_coord.cpp:3129-3144).  voxels_host may be NULL to allocate only (fill later
 * with dscgpu_volume_ptr_dev). */
int dscgpu_ctx_create(dscgpu_ctx **out, int device, const int dims[3], int dtype, const void *voxels_host);
int dscgpu_ctx_destroy(dscgpu_ctx *ctx);
const char *dscgpu_last_error(const dscgpu_ctx *ctx);
/* Launch everything on this cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
int dscgpu_set_stream(dscgpu_ctx *ctx, void *cuda_stream);
int dscgpu_synchronize(dscgpu_ctx *ctx);
/* Device pointer of the voxel volume (dtype as created) and of the z prefix-sum table (FP64). */
void *dscgpu_volume_ptr_dev(dscgpu_ctx *ctx);
/* The tet pass gathers from a bricked copy of the volume (4x2x4-voxel bricks, built once per volume like the sum table).
 * Call this after writing the volume through dscgpu_volume_ptr_dev so that the copy is rebuilt at the next tet pass
 * (fetching the pointer marks it stale as well). */
int dscgpu_volume_modified(dscgpu_ctx *ctx);
/* Replaces the voxels of an existing context (same dims and dtype as at creation): image3d::load on a live object.  The
 * sum table has to be rebuilt afterwards (dscgpu_build_sum_table). */
int dscgpu_volume_upload(dscgpu_ctx *ctx, const void *voxels_host);
double *dscgpu_sum_table_ptr_dev(dscgpu_ctx *ctx);

/* image3d::build_sum_table (DEMO/image3d.cpp:143-177): S[x,y,z] = S[x,y,z-1] + V[x,y,z], bit-exact. */
int dscgpu_build_sum_table(dscgpu_ctx *ctx);
/* Copies the table back (X*Y*Z doubles) -- tests only. */
int dscgpu_get_sum_table(dscgpu_ctx *ctx, double *table_host);

/* Uploads a snapshot (replaces the per-call mesh walks of DEMO/segment_function.cpp:1435-1447,
 * 1626-1634 and image3d.cpp's std::vector<vec3> arguments).  Host pointers are not retained. */
int dscgpu_set_snapshot(dscgpu_ctx *ctx, const dscgpu_snapshot *snap);
/* Zero-copy variant: the library hands out PINNED host arrays sized for the given counts; the host
 * flattens the mesh straight into them (pos4 holds 4 doubles per node: x, y, z, unused) and
 * dscgpu_commit_snapshot uploads them without an intermediate copy. */
typedef struct dscgpu_snapshot_staging {
    double *pos4;
    uint32_t *tet_nodes;
    int32_t *tet_label;
    uint32_t *face_nodes;
    uint32_t *face_tets;
    uint8_t *face_is_boundary;
} dscgpu_snapshot_staging;
int dscgpu_snapshot_staging_get(dscgpu_ctx *ctx, uint32_t n_nodes_buffer, uint32_t n_tets, uint32_t n_faces, dscgpu_snapshot_staging *out);
int dscgpu_commit_snapshot(dscgpu_ctx *ctx);
/* SURVEY 8(f) row 4 -- incremental snapshot.  Between two evaluations of segment() every interface node moves but only the tets
 * that deform()/adapt() touched change (the reference tracks them, src/cache.hpp:82-115); re-sending 52 MB of an otherwise
 * unchanged 2 M-tet snapshot is what bounds the end-to-end step.  This call edits the snapshot resident on the device:
 *   n_pos positions (3 doubles each) for the node keys in node_keys (NULL: keys 0 .. n_pos-1),
 *   n_tet_upd tets replaced in place: tet_index[i] (index into the snapshot's tet arrays) gets tet_nodes[4i..] and tet_label[i],
 *   replace_faces != 0: the interface-face list is replaced as a whole (it is small and depends on every label) and the
 *   node -> face table rebuilt; otherwise it is kept.
 * n_nodes_buffer and n_tets stay what dscgpu_set_snapshot established (a mesh that grew needs a full snapshot).  Out-of-range
 * keys / indices are dropped and reported by the next synchronising call.  Afterwards dscgpu_image_force_eval(ctx, NULL, ...)
 * evaluates the resident snapshot without any upload. */
int dscgpu_update_snapshot(dscgpu_ctx *ctx, uint32_t n_pos, const uint32_t *node_keys, const double *pos, uint32_t n_tet_upd, const uint32_t *tet_index,
                           const uint32_t *tet_nodes, const int32_t *tet_label, int replace_faces, uint32_t n_faces, const uint32_t *face_nodes,
                           const uint32_t *face_tets, const uint8_t *face_is_boundary);
/* Zero-copy variant: PINNED host arrays sized for the given counts (node_keys NULL when with_keys == 0); fill them, then commit.
 * The arrays stay valid -- and may be committed again -- until the next dscgpu_update_staging_get. */
typedef struct dscgpu_update_staging {
    uint32_t *node_keys;
    double *pos;            /* 3 per updated node */
    uint32_t *tet_index;
    uint32_t *tet_nodes;    /* 4 per updated tet */
    int32_t *tet_label;
    uint32_t *face_nodes;   /* the whole new interface-face list, used when dscgpu_update_commit(ctx, 1) */
    uint32_t *face_tets;
    uint8_t *face_is_boundary;
} dscgpu_update_staging;
int dscgpu_update_staging_get(dscgpu_ctx *ctx, uint32_t n_pos, int with_keys, uint32_t n_tet_upd, uint32_t n_faces, dscgpu_update_staging *out);
int dscgpu_update_commit(dscgpu_ctx *ctx, int replace_faces);
/* Multi-GPU: restrict the work of this context to tets [tet_begin,tet_end), interface faces
 * [face_begin,face_end) and (for the gather force kernel) nodes [node_begin,node_end); the
 * snapshot itself stays whole.  Reset by dscgpu_set_snapshot. */
int dscgpu_set_partition(dscgpu_ctx *ctx, uint32_t tet_begin, uint32_t tet_end, uint32_t face_begin, uint32_t face_end,
                         uint32_t node_begin, uint32_t node_end);

/* -- per-tet integration ------------------------------------------ */

/* image3d::get_tetra_intensity / get_energy / get_variation batched over all tets
 * (DEMO/image3d.cpp:351-426; callers DEMO/segment_function.cpp:165-186, 618-643, 850-889).
 * Any output pointer may be NULL.  Tets with label 0 are skipped (outputs -1) unless
 * include_bound != 0.
 *   tet_total[t] = *total_inten, tet_vol[t] = *volume, tet_mean[t] = return value
 *   nc, c[nc]    : candidate intensities; tet_energy[t*nc+k] = get_energy(t, c[k]),
 *                  tet_variation[t*nc+k] = get_variation(t, c[k])
 *   stats        : per-label reduction of tet_total / sum I^2 / tet_vol and mean = total/volume
 */
int dscgpu_tet_integrate(dscgpu_ctx *ctx, int nb_phase, int include_bound, int nc, const double *c, double *tet_total, double *tet_vol,
                         double *tet_mean, double *tet_energy, double *tet_variation, dscgpu_phase_stats *stats);
/* SURVEY 8(f) row 1 -- batched relabel energies: segment_function::get_energy_tetrahedron (DEMO/segment_function.cpp:618-643)
 * for every labelled tet and every candidate label 1..nb_phase, labels left untouched (what relabel_tetrahedra, :845-890,
 * evaluates 2*NB_PHASE times per tet):
 *   energy[t*nb_phase+k] = get_variation(t, mean[k]) + sum over the tet's 4 faces (get_faces(t) order) of alpha*area(face)
 *                          when the label on the other side differs from k+1 (:633-640); -1 for label-0 tets.
 * The tet -> face adjacency is an optional extension of the snapshot, set after dscgpu_set_snapshot:
 *   tet_face_nodes 12*n_tets node keys: per tet its 4 faces, each in get_nodes(f) order (DSC::area, src/DSC.h:3216-3220)
 *   tet_face_tets   8*n_tets tet indices: per face its co-boundary tets in get_tets(f) order, DSCGPU_NO_TET when single
 *                   (such faces are skipped; the reference assumes there are none). */
int dscgpu_set_tet_faces(dscgpu_ctx *ctx, const uint32_t *tet_face_nodes, const uint32_t *tet_face_tets);
int dscgpu_relabel_energies(dscgpu_ctx *ctx, int nb_phase, const double *mean, double alpha, double *energy);

/* Order-sensitive FNV-1a hash of the truncated voxel coordinates of every sample point of every
 * tet, and the sample count: the bit-exactness witness for the sample -> voxel classification. */
int dscgpu_tet_sample_hash(dscgpu_ctx *ctx, uint64_t *tet_hash, int32_t *tet_nsamp);
/* is_point_inside (DEMO/segment_function.cpp:2576-2580): 4-point barycentric voxel-in-tet predicate. */
int dscgpu_points_in_tets(dscgpu_ctx *ctx, uint32_t n, const uint32_t *tet_idx, const double *pts, uint8_t *inside);

/* -- z-ray means and energy ---------------------------------------------- */

/* segment_function::update_average_intensity (DEMO/segment_function.cpp:1600-1764): fills
 * total (= _total_intensities), volume (= _phase_volume, integer-valued voxel counts, bit-exact)
 * and mean.  Needs dscgpu_build_sum_table. */
int dscgpu_zray_means(dscgpu_ctx *ctx, int nb_phase, dscgpu_phase_stats *stats);
/* segment_function::compute_energy (DEMO/segment_function.cpp:2331-2492): the two numbers of its
 * "=== Energy; area = data - area" line.  vertex_bound (n_nodes_buffer bytes) may be NULL, in
 * which case m_vertex_bound is derived from the snapshot as update_vertex_boundary does (:2158-2171). */
int dscgpu_zray_energy(dscgpu_ctx *ctx, int nb_phase, const double *mean, double alpha, const uint8_t *vertex_bound,
                       double *energy_data, double *energy_area);
/* m_vertex_bound as derived from the snapshot (n_nodes_buffer bytes). */
int dscgpu_vertex_bound(dscgpu_ctx *ctx, uint8_t *vertex_bound);

/* -- external force ----------------------------------------------------------- */

enum dscgpu_force_mode {
    DSCGPU_FORCE_GATHER = 0, /* one thread per node walks its incident interface faces in ascending
                                order: same addition order as the reference, no atomics, bit-reproducible */
    DSCGPU_FORCE_ATOMIC = 1  /* one thread per face, FP64 red.global.add scatter */
};
/* segment_function::compute_external_force (DEMO/segment_function.cpp:1425-1473): forces is
 * 3*n_nodes_buffer doubles indexed by raw node key (= _forces). */
int dscgpu_face_forces(dscgpu_ctx *ctx, int nb_phase, const double *mean, int mode, double *forces);

/* SURVEY 8(f) row 2 -- the initialisation pass, segment_function::initialization_discrete_opt (DEMO/segment_function.cpp:165-262):
 *   dscgpu_init_histogram  per-tet mean intensities of all labelled tets (kept on the device) and their 256-bin histogram,
 *                          bin = (int)(mean*256) clamped to 255 (:219-226).  Label-0 tets are skipped; the reference's loop also
 *                          visits the -1 entries of its key-indexed buffer and writes bin -256 for them, out of bounds.
 *   (host)                 thresholds = otsu_muti(hist, NB_PHASE), DEMO/otsu_multi.h:105-129 -- sequential scalar code
 *   dscgpu_init_labels     label = 1 + lower_bound(thresholds, min((int)(mean*255), 255)) for labelled tets, 0 for label-0 tets
 *                          (:245-261); the host applies them with set_label. */
int dscgpu_init_histogram(dscgpu_ctx *ctx, int nb_phase, int32_t hist[256]);
int dscgpu_init_labels(dscgpu_ctx *ctx, int n_thresholds, const int32_t *thresholds, int32_t *tet_label_out);

/* SURVEY 8(f) row 3 -- segment_function::compute_internal_force_2 (DEMO/segment_function.cpp:2197-2232), the curvature force
 * segment() evaluates next to the external force every iteration (:2246): internal_forces is 3*n_nodes_buffer doubles indexed by
 * raw node key (= _internal_forces); get_node_displacement is (forces + alpha*internal_forces)*dt (:581-584).
 *   node_flags    n_nodes_buffer bytes: bit 0 NodeAttributes::is_interface(), bit 1 is_crossing() (is_mesh/attributes.h:84-102).
 *                 is_crossing comes from connected components of the tets around a node (is_mesh/is_mesh.h:523-573): topology,
 *                 kept by the host; reading the attribute is free.
 *   vertex_bound  n_nodes_buffer bytes (m_vertex_bound) or NULL to derive it from the snapshot like dscgpu_vertex_bound.
 * Nodes that are not projected (crossing, or not interface) are bit-identical to the reference.  Projected ones are bit-identical
 * when dscgpu_set_node_face_order handed in the order of DSC::get_normal(node) for this snapshot; otherwise they agree to rounding
 * (the node normal then adds the face normals in ascending face order). */
int dscgpu_internal_forces(dscgpu_ctx *ctx, const uint8_t *node_flags, const uint8_t *vertex_bound, double *internal_forces);

/* The order in which DSC::get_normal(node_key) (src/DSC.h:3138-3159) adds the normals of a node's interface faces: that of
 * ISMesh::get_faces(NodeKey) (is_mesh/is_mesh.h:691-699 -- the node's edges, the faces of each edge, first occurrence; the DSC_CACHE
 * copy get_faces_cache is filled by the same walk, src/DSC.h:243-251).  It is topology of the complex, not derivable from the flat arrays.
 *   node_off  n_nodes_buffer + 1 offsets (node_off[0] = 0, node_off[n_nodes_buffer] = n_entries); nodes the reference does not
 *             project (not is_interface, or is_crossing) may have empty ranges.
 *   face_idx  n_entries indices into the snapshot's interface-face list, the faces of get_faces(node) with is_interface() in that order.
 * Belongs to the snapshot it was set on: dscgpu_set_snapshot, dscgpu_commit_snapshot* and a dscgpu_update_snapshot that replaces the
 * face list drop it.  node_off == NULL drops it.  Returns DSCGPU_ERR_ARG for offsets or indices out of range. */
int dscgpu_set_node_face_order(dscgpu_ctx *ctx, uint32_t n_entries, const uint32_t *node_off, const uint32_t *face_idx);

/* A9b -- the force loop of segment_function::thickenning_surface (DEMO/segment_function.cpp:1990-2051): the external force over the
 * interface faces whose three nodes are not all in m_vertex_bound (segment_function::is_boundary(FaceKey), :2186-2196), and per face the
 * mean and the mean absolute deviation of the scalar magnitudes of its Gauss points (:2030-2043); the reference splits the longest
 * edge of a face when either exceeds get_avg_edge_length()*0.5*0.01 (:2046-2049, host: topology).
 *   vertex_bound   n_nodes_buffer bytes (m_vertex_bound) or NULL to derive it from the snapshot
 *   forces         3*n_nodes_buffer doubles (the reference then adds compute_internal_force_2's, :2053-2054)
 *   face_mean / face_dev / face_evaluated   n_faces entries each (may be NULL); skipped faces report 0, 0, 0
 * Same ordered sums as DSCGPU_FORCE_GATHER: forces and statistics bit-identical to the reference. */
int dscgpu_thicken_forces(dscgpu_ctx *ctx, int nb_phase, const double *mean, const uint8_t *vertex_bound, double *forces, double *face_mean, double *face_dev,
                          uint8_t *face_evaluated);

/* SURVEY 8(f) row 3, second half -- get_node_displacement (DEMO/segment_function.cpp:581-584) and the non-topological part of
 * snapp_boundary (:466-541) for every node at once, so that the forces need not leave the device before set_destination:
 *   node_displacement = (forces + alpha internal_forces) dt, destination = pos + it; nodes of m_vertex_bound are aligned to the faces
 *   of the image domain (macro algin_pos, :59-65, threshold align_threshold = AVG_LENGTH*0.5*0.6); the step is scaled when the largest
 *   displacement of a free node exceeds max_displacement (= get_avg_edge_length()*0.5*0.4, :512-519), bound nodes are clipped (:531-534);
 *   destinations[3 v] = pos + dis * dt_adapt[v] for is_interface() nodes -- what the reference hands to DSC::set_destination (:536) --
 *   and pos for every other node; cur_dis (may be NULL) = _cur_dis.
 * forces / internal_forces / node_flags / vertex_bound: host arrays, or NULL to use what the last dscgpu_image_force_eval /
 * dscgpu_face_forces and dscgpu_internal_forces on this snapshot left on the device.  dt_adapt NULL = 1.0 for every node.
 * Element-wise FP64 in the reference's order and exact maxima: bit-identical to the reference for identical forces. */
typedef struct dscgpu_move_params {
    double dt, alpha, align_threshold, max_displacement;
} dscgpu_move_params;
typedef struct dscgpu_move_stats {
    double max_dis, max_boundary_dis, scale, real_max; /* the numbers snapp_boundary prints; _dt is multiplied by scale (:516) */
} dscgpu_move_stats;
int dscgpu_node_destinations(dscgpu_ctx *ctx, const dscgpu_move_params *params, const double *forces, const double *internal_forces, const uint8_t *node_flags,
                             const uint8_t *vertex_bound, const double *dt_adapt, double *destinations, double *cur_dis, dscgpu_move_stats *stats);

/* -- fused evaluation and device-resident variants ------------------------------------------- */

enum dscgpu_mean_mode { DSCGPU_MEAN_TET = 0, DSCGPU_MEAN_ZRAY = 1 };
/* One full image-force evaluation = snapshot upload + per-tet integration + per-phase means
 * (tet-sampled or z-ray) + external forces + result download: what segment() needs from
 * update_average_intensity + compute_external_force (DEMO/segment_function.cpp:2244-2247). */
int dscgpu_image_force_eval(dscgpu_ctx *ctx, const dscgpu_snapshot *snap, int nb_phase, int mean_mode, int force_mode,
                            dscgpu_phase_stats *stats, double *forces);

/* After dscgpu_image_force_eval: the forces as they arrived in the context's PINNED host buffer (3*n_nodes_buffer doubles,
 * valid until the next call on the context).  Passing forces == NULL to dscgpu_image_force_eval and reading them here
 * avoids one host-side copy per evaluation. */
const double *dscgpu_forces_pinned(dscgpu_ctx *ctx, uint32_t *n_nodes_buffer);
/* Compact read-back: with dscgpu_set_compact_forces(ctx, 1), dscgpu_image_force_eval(..., forces = NULL) brings back only the rows of
 * the nodes that have interface faces -- every other row of _forces is zero (DEMO/segment_function.cpp:1431: the vector starts at
 * zero and only nodes of interface faces are added to) -- about one node in seven on a DSC mesh.  dscgpu_forces_compact_pinned
 * gives the pinned result: rows[3 i .. 3 i + 2] is the force of node node_keys[i], keys ascending. */
int dscgpu_set_compact_forces(dscgpu_ctx *ctx, int on);
int dscgpu_forces_compact_pinned(dscgpu_ctx *ctx, uint32_t *n_rows, const uint32_t **node_keys, const double **rows);
/* The same read-back for a force array that lives on the device (the output of dscgpu_face_forces_dev / _allreduce_dev): gathers the
 * rows of the nodes with interface faces, copies them and their keys to the context's pinned buffers, synchronises. */
int dscgpu_read_forces_compact(dscgpu_ctx *ctx, const double *d_forces, uint32_t *n_rows, const uint32_t **node_keys, const double **rows);

/* Device-resident pieces for multi-GPU composition (NCCL all-reduce happens between them, on
 * the caller's stream set with dscgpu_set_stream):
 *   d_sums   3*nb_phase doubles: total[nb], sumsq[nb], volume[nb] of this context's partition
 *   d_mean   nb_phase doubles
 *   d_forces 3*n_nodes_buffer doubles, overwritten */
int dscgpu_tet_phase_sums_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums);
int dscgpu_zray_phase_sums_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums);
int dscgpu_means_from_sums_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_sums, double *d_mean);
/* Starts the part of the external force that does not depend on the phase means -- face normals, areas, Gauss sets and the
 * trilinear sample of every Gauss point -- on the context's second stream, so that it runs next to the tet pass that produces
 * the means; the next DSCGPU_FORCE_GATHER force call (dscgpu_face_forces[_dev]) joins it and only forms and adds the vectors.
 * Optional: without it the force call computes that part itself.  Call it after the snapshot (and partition) is in place. */
int dscgpu_face_geometry_async(dscgpu_ctx *ctx);
/* With on = 1, dscgpu_tet_phase_sums_dev and the tet pass of dscgpu_image_force_eval start dscgpu_face_geometry_async themselves, right
 * after their main kernel: the geometry half of the force pass then runs next to the short kernels that end the tet pass.  Use it
 * when a force call follows every tet pass (the per-iteration pattern); results are unchanged. */
int dscgpu_set_overlap_geometry(dscgpu_ctx *ctx, int on);
int dscgpu_face_forces_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_mean, int mode, double *d_forces);

/* -- multi-GPU exchange over NVLink peer memory ------------------------------------------------------------------------ */

/* One process per GPU.  The two exchanges of an evaluation (3*nb_phase per-phase sums after the tet pass; the force rows of
 * the nodes that have interface faces) run as kernels that read the peers' windows directly (csrc/dsc_comm.cuh) instead of
 * two NCCL all-reduces: same results on every rank (peers are added in rank order), no pack / unpack kernels, and the whole
 * evaluation can be captured in one CUDA graph.
 *   dscgpu_comm_create   allocates this rank's window for snapshots of up to cap_rows nodes (n_nodes_buffer) and returns its
 *                        64-byte CUDA IPC handle; world <= 8
 *   dscgpu_comm_connect  maps the peers' windows: handles = world x 64 bytes in rank order (exchange them with any host-side
 *                        all-gather); every rank must have created its window before any rank runs an exchange
 *   dscgpu_comm_connect_local   the same for contexts that live in ONE process (tests): windows[k] = dscgpu_comm_window_dev
 *                        of rank k's context
 *   dscgpu_comm_error    synchronises and reports a timed-out exchange (a peer that never arrived within 2 s) */
int dscgpu_comm_create(dscgpu_ctx *ctx, int rank, int world, uint32_t cap_rows, unsigned char handle_out[64]);
/* The same with a snapshot region of snapshot_bytes in the window, for dscgpu_commit_snapshot_sharded: size it for the largest
 * snapshot, 32 n_nodes_buffer + 20 n_tets + 21 n_faces + 2 KB. */
int dscgpu_comm_create_ex(dscgpu_ctx *ctx, int rank, int world, uint32_t cap_rows, uint64_t snapshot_bytes, unsigned char handle_out[64]);
/* dscgpu_commit_snapshot when every rank holds the same snapshot in its pinned staging arrays (dscgpu_snapshot_staging_get): rank r
 * sends only its 1/world share of the bytes over PCIe, the shares travel to the peers' windows over NVLink, and every rank ends
 * up with the whole snapshot.  Collective: every rank must call it, with identical staging contents.  Resets the partition. */
int dscgpu_commit_snapshot_sharded(dscgpu_ctx *ctx);
int dscgpu_comm_connect(dscgpu_ctx *ctx, const unsigned char *handles);
int dscgpu_comm_connect_local(dscgpu_ctx *ctx, void *const *windows);
void *dscgpu_comm_window_dev(dscgpu_ctx *ctx);
int dscgpu_comm_error(dscgpu_ctx *ctx);
/* dscgpu_tet_phase_sums_dev over this context's tet partition fused with the all-reduce: d_sums (3*nb_phase) and d_mean
 * (nb_phase, may be NULL) hold the totals over ALL ranks.  Must be followed by dscgpu_face_forces_allreduce_dev on every rank
 * (one exchange epoch per evaluation). */
int dscgpu_tet_phase_sums_allreduce_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums, double *d_mean);
/* External forces of this context's face partition (atomic scatter into compact rows) fused with the all-reduce:
 * d_forces (3*n_nodes_buffer) holds the forces of ALL ranks' faces; 1e-6 parity like DSCGPU_FORCE_ATOMIC. */
int dscgpu_face_forces_allreduce_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_mean, double *d_forces);

/* -- instrumentation -------------------------------------------------------------------------- */

enum dscgpu_timer {
    DSCGPU_T_TET = 0, DSCGPU_T_PHASE_FINAL = 1, DSCGPU_T_FORCE = 2, DSCGPU_T_ZRAY = 3, DSCGPU_T_SUMTABLE = 4,
    DSCGPU_T_H2D = 5, DSCGPU_T_D2H = 6, DSCGPU_T_COUNT = 7
};
/* When enabled, kernels are bracketed by CUDA events on the launching stream; dscgpu_get_timer
 * synchronizes and returns the duration of the last launch of that class in ms. */
int dscgpu_enable_timers(dscgpu_ctx *ctx, int on);
int dscgpu_get_timer(dscgpu_ctx *ctx, int which, float *ms);
/* Number of kernels this library has launched on the context since creation. */
uint64_t dscgpu_launch_count(const dscgpu_ctx *ctx);
/* Tunables: lanes per tet of the integration kernel (1, 8 or 32; 0 = choose from the mesh). */
int dscgpu_set_lanes_per_tet(dscgpu_ctx *ctx, int lanes);
/* Implementation of the per-tet integration kernel (every one puts every sample in the reference's voxel; sums agree to rounding):
 *   0             v1: G lanes per tet (dscgpu_set_lanes_per_tet), the reference's FP64 sequence for every sample
 *   bit 7 (128)   w2: warp-level two-phase kernel, origin-relative lattice coordinates, FP32 filter with exact FP64 fallback
 *   bit 12 (4096) with bit 7: keep the general float widening (off: scaled widening when the volume allows it)
 *   bit 14 (16384) with bit 7: full batches without row clamps, per-tet weight from phase A, on the sums-only path
 *   bit 17 (131072) staged kernel: lane per tet, the voxel box of each 32-tet tile brought into shared memory by TMA, per-label sums
 *                 in 64-bit fixed point (independent of the order of the tiles and of the partition); volumes it cannot describe
 *                 (FP64 voxels, row pitch not a multiple of 16 bytes, float volumes with negative or non-finite voxels) fall
 *                 through to the kernel the remaining bits select
 * The environment variable DSCGPU_TET_VARIANT overrides the default at context creation; DESIGN.md section 9 has the measurements,
 * the branch proto/round1-variants the forms that were measured and dropped.
 * Per-candidate outputs (tet_energy / tet_variation) always use the v1 kernel. */
int dscgpu_set_tet_variant(dscgpu_ctx *ctx, int variant);
int dscgpu_get_tet_variant(const dscgpu_ctx *ctx);
/* The per-iteration entry points (dscgpu_image_force_eval, dscgpu_tet_phase_sums_dev and the *_allreduce_dev forms) fill the
 * per-label second moment (sum of I^2 v/n^3, DEMO/image3d.cpp:384-404 with c = 0) only when asked: 0 (default) leaves it zero. */
int dscgpu_set_hot_want_sq(dscgpu_ctx *ctx, int on);
/* Algorithmic sample count of the last tet pass (sum of n^3 over processed tets). */
uint64_t dscgpu_last_sample_count(dscgpu_ctx *ctx);
uint64_t dscgpu_last_gauss_count(dscgpu_ctx *ctx);

/* libstdc++ std::sort restated for the z-ray hits (introsort, threshold 16, median-of-3), run on
 * the HOST instance of the same __host__ __device__ code the kernel uses -- lets CPU-only tests
 * fuzz it against the real std::sort. */
int dscgpu_host_sort_hits(int n, int32_t *z, uint8_t *b_in);

#ifdef __cplusplus
}
#endif
#endif /* DSCGPU_H */
