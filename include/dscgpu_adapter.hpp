// dscgpu_adapter.hpp -- C++ host adapter between a live reference mesh and the C ABI (include/dscgpu.h).
//
// Header-only; compile it in a translation unit that already sees the reference's headers
// (DEMO/segment_function.h -> DSC.h, is_mesh.h, image3d.h).  It does what SURVEY.md 8(b) item 6 asks for:
//   * snapshot_from_dsc(dsc, snap): flattens the is_mesh/DSC complex through its public read-only getters, in ascending
//     key order, into the SoA arrays of dscgpu_snapshot -- the same walk the reference's own loops do
//     (tetrahedra_begin()/faces_begin(), DEMO/segment_function.cpp:1435, 1626; get_nodes/get_tets/get_label/get_pos,
//     is_mesh/is_mesh.h:250-253, 625-679);
//   * ImageEnergy: owns a dscgpu context for one image3d and offers drop-in bodies for
//       segment_function::update_average_intensity   (DEMO/segment_function.cpp:1600-1764)
//       segment_function::compute_external_force     (DEMO/segment_function.cpp:1425-1473)
//       segment_function::compute_energy             (DEMO/segment_function.cpp:2331-2492)
//       segment_function::compute_internal_force_2   (DEMO/segment_function.cpp:2197-2232)
//     that write the reference's public result members (_mean_intensities, _total_intensities, _phase_volume, _forces,
//     _internal_forces).
// INTEGRATION.md shows the three-line patch to DEMO/segment_function.cpp that routes segment() through it.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "dscgpu.h"

namespace dscgpu {

struct Snapshot {
    std::vector<double> pos;            // 3 * get_no_nodes_buffer()
    std::vector<uint32_t> tet_key;      // raw keys of the valid tets, ascending
    std::vector<uint32_t> tet_nodes;    // 4 per tet, get_nodes(t) order
    std::vector<int32_t> tet_label;
    std::vector<uint32_t> face_key;     // raw keys of the interface faces, ascending
    std::vector<uint32_t> face_nodes;   // 3 per face, get_nodes(f) order
    std::vector<uint32_t> face_tets;    // 2 per face: indices into the tet arrays, get_tets(f) order
    std::vector<uint8_t> face_is_boundary;
    std::vector<uint8_t> node_flags;    // per node of the buffer: bit 0 is_interface(), bit 1 is_crossing() (is_mesh/attributes.h:84-102)

    void swap(Snapshot &o)
    {
        pos.swap(o.pos); tet_key.swap(o.tet_key); tet_nodes.swap(o.tet_nodes); tet_label.swap(o.tet_label); face_key.swap(o.face_key);
        face_nodes.swap(o.face_nodes); face_tets.swap(o.face_tets); face_is_boundary.swap(o.face_is_boundary); node_flags.swap(o.node_flags);
    }

    dscgpu_snapshot view() const
    {
        dscgpu_snapshot s;
        s.n_nodes_buffer = (uint32_t)(pos.size() / 3);
        s.pos = pos.data();
        s.n_tets = (uint32_t)tet_label.size();
        s.tet_nodes = tet_nodes.data();
        s.tet_label = tet_label.data();
        s.n_faces = (uint32_t)face_is_boundary.size();
        s.face_nodes = face_nodes.data();
        s.face_tets = face_tets.data();
        s.face_is_boundary = face_is_boundary.data();
        return s;
    }
};

// DSC is DSC::DeformableSimplicialComplex<> (or any ISMesh with the same getters).
// key_indexed: the tet arrays have one entry per slot of the tet buffer (get_no_tets_buffer(), index = raw key; free slots
// carry DSCGPU_LABEL_FREE and are skipped by every kernel) instead of one per valid tet.  deform() collapses and splits simplices in
// every iteration, so a dense list shifts all the time; with key-indexed arrays an unchanged tet keeps its place and
// ImageEnergy::refresh can send differences only.
struct WalkTimes { double nodes_ms = 0, tets_ms = 0, faces_ms = 0; }; // where snapshot_from_dsc spends its time (oracle/walk_timing.cpp)

template <class DSC> void snapshot_from_dsc(DSC &dsc, Snapshot &s, bool key_indexed = false, WalkTimes *times = nullptr)
{
    auto clock_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = clock_ms();
    const unsigned nnb = dsc.get_no_nodes_buffer();
    s.pos.assign(3 * (size_t)nnb, 0.0);
    s.node_flags.assign(nnb, 0);
    for (auto nit = dsc.nodes_begin(); nit != dsc.nodes_end(); nit++) {
        const unsigned k = nit.key();
        const auto p = dsc.get_pos(nit.key());
        s.pos[3 * (size_t)k] = p[0]; s.pos[3 * (size_t)k + 1] = p[1]; s.pos[3 * (size_t)k + 2] = p[2];
        s.node_flags[k] = (uint8_t)((nit->is_interface() ? 1 : 0) | (nit->is_crossing() ? 2 : 0));
    }
    const double t_nodes = clock_ms();
    std::vector<int32_t> tet_index(dsc.get_no_tets_buffer(), -1);
    s.tet_key.clear(); s.tet_nodes.clear(); s.tet_label.clear();
    if (key_indexed) {
        const size_t ntb = dsc.get_no_tets_buffer();
        s.tet_key.resize(ntb);
        s.tet_nodes.assign(4 * ntb, 0u);
        s.tet_label.assign(ntb, DSCGPU_LABEL_FREE);
        for (size_t k = 0; k < ntb; k++) s.tet_key[k] = (uint32_t)k;
        for (auto tit = dsc.tetrahedra_begin(); tit != dsc.tetrahedra_end(); tit++) {
            const size_t k = (size_t)(unsigned)tit.key();
            tet_index[k] = (int32_t)k;
            // get_nodes(t) builds the node set by set unions on every call (is_mesh/is_mesh.h:673-679, ~0.7 us per tet); with DSC_CACHE
            // the complex keeps it per tet and drops it when the tet changes (src/cache.hpp:107-111, src/DSC.h:290-299): same order
#ifdef DSC_CACHE
            const auto &nids = *dsc.get_nodes_cache(tit.key());
#else
            const auto nids = dsc.get_nodes(tit.key());
#endif
            for (int c = 0; c < 4; c++) s.tet_nodes[4 * k + c] = nids[c];
            s.tet_label[k] = dsc.get_label(tit.key());
        }
    } else
    for (auto tit = dsc.tetrahedra_begin(); tit != dsc.tetrahedra_end(); tit++) {
        tet_index[tit.key()] = (int32_t)s.tet_key.size();
        s.tet_key.push_back(tit.key());
#ifdef DSC_CACHE
        const auto &nids = *dsc.get_nodes_cache(tit.key());
#else
        const auto nids = dsc.get_nodes(tit.key());
#endif
        for (int k = 0; k < 4; k++) s.tet_nodes.push_back(nids[k]);
        s.tet_label.push_back(dsc.get_label(tit.key()));
    }
    const double t_tets = clock_ms();
    s.face_key.clear(); s.face_nodes.clear(); s.face_tets.clear(); s.face_is_boundary.clear();
    for (auto fit = dsc.faces_begin(); fit != dsc.faces_end(); fit++) {
        if (!fit->is_interface()) continue;
        s.face_key.push_back(fit.key());
        // (get_nodes(f) unions the node sets of two edges on every call, is_mesh/is_mesh.h:665-671; the complex's own per-face copy holds the same set)
#ifdef DSC_CACHE
        const auto &nids = *dsc.get_nodes_cache(fit.key());
#else
        const auto nids = dsc.get_nodes(fit.key());
#endif
        for (int k = 0; k < 3; k++) s.face_nodes.push_back(nids[k]);
        const auto &tids = dsc.get_tets(fit.key());
        s.face_tets.push_back((uint32_t)tet_index[tids[0]]);
        s.face_tets.push_back(tids.size() > 1 ? (uint32_t)tet_index[tids[1]] : DSCGPU_NO_TET);
        s.face_is_boundary.push_back(fit->is_boundary() ? 1 : 0);
    }
    if (times) { times->nodes_ms = t_nodes - t_start; times->tets_ms = t_tets - t_nodes; times->faces_ms = clock_ms() - t_tets; }
}

class Error : public std::runtime_error {
public:
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

// IMG is the reference's image3d; SEG its segment_function.
template <class IMG, class SEG> class ImageEnergy {
public:
    // Uploads img._voxels and builds the z prefix-sum table on the device.  image3d::load fills _voxels with byte / 255.0
    // (DEMO/image3d.cpp:127): when every voxel is exactly such a value the volume goes up as bytes -- the device decodes q / 255.0 to
    // the same doubles, at 1/8 of the memory and gather traffic; anything else goes up as the FP64 values themselves.
    explicit ImageEnergy(IMG &img, int device = 0, int force_mode = DSCGPU_FORCE_GATHER) : force_mode_(force_mode)
    {
        const int dims[3] = {img._dim[0], img._dim[1], img._dim[2]};
        const size_t n = img._voxels.size();
        std::vector<uint8_t> bytes(n);
        bool as_bytes = true;
        for (size_t i = 0; i < n && as_bytes; i++) {
            const double v = img._voxels[i], q = std::nearbyint(v * 255.0);
            if (!(q >= 0.0 && q <= 255.0) || q / 255.0 != v) as_bytes = false;
            else bytes[i] = (uint8_t)q;
        }
        volume_dtype_ = as_bytes ? DSCGPU_U8 : DSCGPU_F64;
        check(dscgpu_ctx_create(&ctx_, device, dims, volume_dtype_, as_bytes ? static_cast<const void *>(bytes.data()) : static_cast<const void *>(img._voxels.data())),
              "dscgpu_ctx_create");
        check(dscgpu_build_sum_table(ctx_), "dscgpu_build_sum_table");
    }
    int volume_dtype() const { return volume_dtype_; }
    ~ImageEnergy() { dscgpu_ctx_destroy(ctx_); }
    ImageEnergy(const ImageEnergy &) = delete;
    ImageEnergy &operator=(const ImageEnergy &) = delete;

    // Call after anything that changed the mesh (deform, relabel, adapt) and after update_vertex_boundary().
    // When the buffers kept their sizes (no key was added: the usual case between two deformations) only what differs from
    // the snapshot resident on the device is sent -- moved nodes, tets whose nodes or label changed, and the (small)
    // interface-face list -- through dscgpu_update_snapshot; otherwise the whole snapshot.
    void refresh(SEG &seg)
    {
        const auto t0 = std::chrono::steady_clock::now();
        struct Stop {
            ImageEnergy *self; std::chrono::steady_clock::time_point t0;
            ~Stop() { self->last_refresh_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
        } stop{this, t0};
        Snapshot &next = scratch_; // (the arrays of the snapshot before last: their capacity is reused)
        snapshot_from_dsc(*seg._dsc, next, /*key_indexed=*/true);
        last_flatten_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const bool same_shape = have_snapshot_ && next.pos.size() == snap_.pos.size() && next.tet_label.size() == snap_.tet_label.size();
        if (!same_shape) {
            snap_.swap(next);
            dscgpu_snapshot v = snap_.view();
            check(dscgpu_set_snapshot(ctx_, &v), "dscgpu_set_snapshot");
            have_snapshot_ = true;
            last_refresh_incremental_ = false;
            return;
        }
        std::vector<uint32_t> keys, tidx, tnodes;
        std::vector<double> pos;
        std::vector<int32_t> tlabel;
        const size_t nn = next.pos.size() / 3, nt = next.tet_label.size();
        for (size_t k = 0; k < nn; k++)
            if (next.pos[3 * k] != snap_.pos[3 * k] || next.pos[3 * k + 1] != snap_.pos[3 * k + 1] || next.pos[3 * k + 2] != snap_.pos[3 * k + 2]) {
                keys.push_back((uint32_t)k);
                pos.insert(pos.end(), next.pos.begin() + 3 * k, next.pos.begin() + 3 * k + 3);
            }
        for (size_t t = 0; t < nt; t++)
            if (next.tet_label[t] != snap_.tet_label[t] || !std::equal(next.tet_nodes.begin() + 4 * t, next.tet_nodes.begin() + 4 * t + 4, snap_.tet_nodes.begin() + 4 * t)) {
                tidx.push_back((uint32_t)t);
                tnodes.insert(tnodes.end(), next.tet_nodes.begin() + 4 * t, next.tet_nodes.begin() + 4 * t + 4);
                tlabel.push_back(next.tet_label[t]);
            }
        check(dscgpu_update_snapshot(ctx_, (uint32_t)keys.size(), keys.data(), pos.data(), (uint32_t)tidx.size(), tidx.data(), tnodes.data(), tlabel.data(), 1,
                                     (uint32_t)next.face_is_boundary.size(), next.face_nodes.data(), next.face_tets.data(), next.face_is_boundary.data()),
              "dscgpu_update_snapshot");
        snap_.swap(next);
        last_refresh_incremental_ = true;
        last_moved_nodes_ = keys.size();
        last_changed_tets_ = tidx.size();
    }
    bool last_refresh_incremental() const { return last_refresh_incremental_; }
    size_t last_moved_nodes() const { return last_moved_nodes_; }
    size_t last_changed_tets() const { return last_changed_tets_; }
    double last_flatten_ms() const { return last_flatten_ms_; } // the is_mesh walk of the last refresh
    double last_refresh_ms() const { return last_refresh_ms_; } // walk + diff + upload (the upload is asynchronous: it overlaps what follows)

    // Body of segment_function::update_average_intensity.
    void update_average_intensity(SEG &seg)
    {
        dscgpu_phase_stats st;
        check(dscgpu_zray_means(ctx_, seg.NB_PHASE, &st), "dscgpu_zray_means");
        seg._mean_intensities.assign(st.mean, st.mean + seg.NB_PHASE);
        seg._total_intensities.assign(st.total, st.total + seg.NB_PHASE);
        seg._phase_volume.assign(st.volume, st.volume + seg.NB_PHASE);
    }

    // Body of segment_function::compute_external_force; _forces is sized get_no_nodes_buffer() and indexed by node key.
    void compute_external_force(SEG &seg)
    {
        const size_t nnb = snap_.pos.size() / 3;
        forces_.resize(3 * nnb);
        check(dscgpu_face_forces(ctx_, seg.NB_PHASE, seg._mean_intensities.data(), force_mode_, forces_.data()), "dscgpu_face_forces");
        seg._forces.resize(nnb);
        for (size_t i = 0; i < nnb; i++) {
            seg._forces[i][0] = forces_[3 * i]; seg._forces[i][1] = forces_[3 * i + 1]; seg._forces[i][2] = forces_[3 * i + 2];
        }
    }

    // Second half of segment_function::update_vertex_boundary (DEMO/segment_function.cpp:2160-2172): m_vertex_bound = nodes of the interface
    // faces that touch a BOUND_LABEL tet, from the snapshot on the device (the first half -- set_label(BOUND_LABEL) on the tets around
    // boundary nodes, :2149-2157 -- edits the complex and stays with the host).  Call refresh() first.
    void update_vertex_boundary_flags(SEG &seg)
    {
        const size_t nnb = snap_.pos.size() / 3;
        std::vector<uint8_t> vb(nnb, 0);
        check(dscgpu_vertex_bound(ctx_, vb.data()), "dscgpu_vertex_bound");
        seg.m_vertex_bound.assign(nnb, false);
        for (size_t i = 0; i < nnb; i++) if (vb[i]) seg.m_vertex_bound[i] = true;
    }

    // Body of segment_function::initialization_discrete_opt (DEMO/segment_function.cpp:143-262): per-tet means and their
    // histogram on the device, the reference's own otsu_muti (pass it in: DEMO/otsu_multi.h defines it in a header) on the
    // host, the label rule on the device, set_label on the host.  Call refresh() first (labels: boundary layer = 0).
    template <class OTSU> void initialization_discrete_opt(SEG &seg, OTSU otsu_muti)
    {
        std::vector<int32_t> hist(256, 0);
        check(dscgpu_init_histogram(ctx_, seg.NB_PHASE, hist.data()), "dscgpu_init_histogram");
        const std::vector<int> thres = otsu_muti(std::vector<int>(hist.begin(), hist.end()), seg.NB_PHASE);
        std::vector<int32_t> t32(thres.begin(), thres.end()), labels(snap_.tet_label.size(), 0);
        check(dscgpu_init_labels(ctx_, (int)t32.size(), t32.data(), labels.data()), "dscgpu_init_labels");
        for (size_t i = 0; i < labels.size(); i++) // (free slots of a key-indexed snapshot carry a negative label: never touched)
            if (labels[i] > 0) seg._dsc->set_label(typename std::decay<decltype(seg._dsc->tetrahedra_begin().key())>::type(snap_.tet_key[i]), labels[i]);
    }

    // The order in which DSC::get_normal(node) adds the face normals (src/DSC.h:3138-3159) for the nodes compute_internal_force_2
    // projects: the interface faces of get_faces(node) -- with DSC_CACHE the complex's own per-node copy, which get_normal reads and the
    // reference's segment() has filled anyway -- as indices into the snapshot's face list.  Topology, so it is walked here and handed
    // to the device; with it the projected forces are bit-identical to the reference's, without it they agree to 1e-12.
    void send_node_face_order(SEG &seg)
    {
        const auto t0 = std::chrono::steady_clock::now();
        auto &dsc = *seg._dsc;
        const size_t nnb = snap_.pos.size() / 3;
        std::vector<int32_t> face_index(dsc.get_no_faces_buffer(), -1);
        for (size_t i = 0; i < snap_.face_key.size(); i++) face_index[snap_.face_key[i]] = (int32_t)i;
        std::vector<uint32_t> off(nnb + 1, 0), idx;
        for (auto nit = dsc.nodes_begin(); nit != dsc.nodes_end(); nit++) {
            const unsigned k = (unsigned)nit.key();
            if (!(snap_.node_flags[k] & 1) || (snap_.node_flags[k] & 2)) continue;
#ifdef DSC_CACHE
            const auto &fids = *dsc.get_faces_cache(nit.key());
#else
            const auto fids = dsc.get_faces(nit.key());
#endif
            for (auto f : fids) {
                const int32_t i = face_index[(unsigned)f];
                if (i >= 0) { idx.push_back((uint32_t)i); off[k + 1]++; } // face_index holds exactly the is_interface() faces
            }
        }
        for (size_t k = 0; k < nnb; k++) off[k + 1] += off[k];
        check(dscgpu_set_node_face_order(ctx_, (uint32_t)idx.size(), off.data(), idx.data()), "dscgpu_set_node_face_order");
        last_face_order_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    double last_face_order_ms() const { return last_face_order_ms_; }

    // Body of segment_function::compute_internal_force_2; m_vertex_bound is read from seg (update_vertex_boundary ran before).
    // exact_normal_order: walk get_faces(node) for the projected nodes first (send_node_face_order): bit-identical forces.
    void compute_internal_force_2(SEG &seg, bool exact_normal_order = true)
    {
        const size_t nnb = snap_.pos.size() / 3;
        if (exact_normal_order) send_node_face_order(seg);
        else check(dscgpu_set_node_face_order(ctx_, 0, nullptr, nullptr), "dscgpu_set_node_face_order");
        forces_.resize(3 * nnb);
        std::vector<uint8_t> vb(nnb, 0);
        for (size_t i = 0; i < nnb && i < seg.m_vertex_bound.size(); i++) vb[i] = seg.m_vertex_bound[i] ? 1 : 0;
        check(dscgpu_internal_forces(ctx_, snap_.node_flags.data(), vb.data(), forces_.data()), "dscgpu_internal_forces");
        seg._internal_forces.resize(nnb);
        for (size_t i = 0; i < nnb; i++) {
            seg._internal_forces[i][0] = forces_[3 * i]; seg._internal_forces[i][1] = forces_[3 * i + 1]; seg._internal_forces[i][2] = forces_[3 * i + 2];
        }
    }

    // Body of the non-topological part of segment_function::snapp_boundary (DEMO/segment_function.cpp:466-541): displacements from the
    // forces, alignment of the m_vertex_bound nodes to the faces of the image, step scaling, set_destination.  The external forces are
    // the ones the last compute_external_force left on the device, the internal ones those of compute_internal_force_2 (call both
    // first, as segment() does, :2246-2247).  The two thresholds are function-local statics in the reference, fixed at the first call.
    dscgpu_move_stats snapp_boundary(SEG &seg)
    {
        if (!move_init_) {
            move_threshold_ = seg._dsc->AVG_LENGTH * 0.5 * 0.6;
            move_max_displacement_ = seg._dsc->get_avg_edge_length() * 0.5 * 0.4;
            move_init_ = true;
        }
        const size_t nnb = snap_.pos.size() / 3;
        dscgpu_move_params mp;
        mp.dt = seg._dt; mp.alpha = seg.m_alpha; mp.align_threshold = move_threshold_; mp.max_displacement = move_max_displacement_;
        seg._dt_adapt.resize(nnb, 1.0);
        dest_.resize(3 * nnb); forces_.resize(3 * nnb);
        dscgpu_move_stats st;
        check(dscgpu_node_destinations(ctx_, &mp, nullptr, nullptr, nullptr, nullptr, seg._dt_adapt.data(), dest_.data(), forces_.data(), &st), "dscgpu_node_destinations");
        seg._dt *= st.scale; // :516
        seg._previous_dis.resize(nnb, typename std::decay<decltype(seg._cur_dis[0])>::type(0));
        seg._cur_dis.resize(nnb, typename std::decay<decltype(seg._cur_dis[0])>::type(0));
        for (auto nit = seg._dsc->nodes_begin(); nit != seg._dsc->nodes_end(); nit++) {
            if (!nit->is_interface()) continue;
            const size_t k = (size_t)(unsigned)nit.key();
            typename std::decay<decltype(seg._cur_dis[0])>::type d(dest_[3 * k], dest_[3 * k + 1], dest_[3 * k + 2]);
            seg._dsc->set_destination(nit.key(), d);
            seg._cur_dis[k][0] = forces_[3 * k]; seg._cur_dis[k][1] = forces_[3 * k + 1]; seg._cur_dis[k][2] = forces_[3 * k + 2];
        }
        // the per-node step factors for the next call (:545-578): host state (the is_clean marks live in the complex's cache), a few
        // operations per interface node -- restated here so that the body is complete
        typedef typename std::decay<decltype(seg._dsc->nodes_begin().key())>::type node_key_t;
        static const double thres_minus = std::cos(M_PI / 180.0 * 150), thres_positive = std::cos(M_PI / 180.0 * 30);
        for (size_t i = 0; i < seg._dt_adapt.size(); i++) {
            node_key_t nk((unsigned)i);
            if (!(seg._dsc->exists(nk) && seg._dsc->get(nk).is_interface() && (!seg.m_vertex_bound[i] || (seg.m_vertex_bound[i] && seg._dsc->get(nk).is_crossing())))) continue;
            if (seg._cur_dis[i].length() > 0.0001) {
                if (!seg._dsc->cache.is_clean[i]) {
                    seg._dsc->cache.is_clean[i] = new bool(true);
                    seg._dt_adapt[i] = 1.0;
                }
                seg._cur_dis[i].normalize();
                const double cos_sign = dot(seg._cur_dis[i], seg._previous_dis[i]);
                if (cos_sign < thres_minus) seg._dt_adapt[i] = std::max(seg._dt_adapt[i] * 0.9, 0.1);
                if (cos_sign > thres_positive) seg._dt_adapt[i] = std::min(seg._dt_adapt[i] * 1.1, 3.0);
                seg._previous_dis[i] = seg._cur_dis[i];
            }
        }
        return st;
    }

    // The two numbers of compute_energy's "=== Energy; area = data - area" line.
    void compute_energy(SEG &seg, double *data, double *area)
    {
        check(dscgpu_zray_energy(ctx_, seg.NB_PHASE, seg._mean_intensities.data(), seg.m_alpha, nullptr, data, area), "dscgpu_zray_energy");
    }

    const Snapshot &snapshot() const { return snap_; }
    dscgpu_ctx *ctx() { return ctx_; }

private:
    void check(int r, const char *what)
    {
        if (r != DSCGPU_OK) throw Error(std::string(what) + ": " + dscgpu_last_error(ctx_));
    }
    dscgpu_ctx *ctx_ = nullptr;
    Snapshot snap_;
    bool have_snapshot_ = false, last_refresh_incremental_ = false;
    size_t last_moved_nodes_ = 0, last_changed_tets_ = 0;
    double last_flatten_ms_ = 0, last_refresh_ms_ = 0, last_face_order_ms_ = 0;
    Snapshot scratch_;
    int volume_dtype_ = DSCGPU_F64;
    std::vector<double> forces_, dest_;
    bool move_init_ = false;
    double move_threshold_ = 0, move_max_displacement_ = 0;
    int force_mode_;
};

} // namespace dscgpu
