// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless driver around the UNMODIFIED reference sources under /root/reference
// (DEMO/image3d.cpp, DEMO/tet_dis_coord.cpp, DEMO/segment_function.cpp, src/DSC.cpp, ...;
// see oracle/Makefile).  It restates only the two GUI-side helpers that cannot be
// compiled without OpenGL -- UI::init_dsc (DEMO/user_interface.cpp:550-630) and
// UI::set_dsc_boundary_layer (DEMO/user_interface.cpp:441-462) -- and then calls the
// reference's own segment_function / image3d members:
//     initialization_discrete_opt, update_vertex_boundary, update_average_intensity,
//     compute_external_force, compute_energy, relabel-energy helpers, segment().
// Every quantity written out below is produced by reference code; this file only
// flattens the is_mesh structures (through the public getters, in key order) and
// writes raw arrays that tests/golden/make_golden.py packs into fixtures.
//
// modes:
//   dump : build mesh, label, run the hot path (optionally after K segment() iterations),
//          write snapshot + expected outputs to --out DIR
//   time : as dump but only prints wall-clock timings of the hot-path members (JSON line)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <random>
#include <memory>
#include <algorithm>

#include "segment_function.h"
#include "tet_dis_coord.hpp"
#include "gaussian_points.h"

// DEMO/otsu_multi.h defines (not declares) its functions; the definition comes from the reference's segment_function.o
std::vector<int> otsu_muti(std::vector<int> histogram_in, int classes);

// globals the reference expects the (GUI) driver to define
// (DEMO/demo.cpp:22,27 and DEMO/user_interface.cpp:187)
double noise_level = 0;
int num_images = 0;
bool arg_b_build_table_origin = true;

typedef DSC::DeformableSimplicialComplex<> dsc_class;

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Args {
    std::string mode = "dump";
    std::string save_dsc; // dump mode: also write the mesh with the reference's own writer (is_mesh::export_tet_mesh, is_mesh/mesh_io.cpp:298-314)
    std::string volume_path;
    std::string vol_dtype = "u8"; // u8 | u16 | f32 | f64
    int dims[3] = {0, 0, 0};
    double edge = 6;
    int nb_phase = 3;
    double alpha = 0.001;
    std::string label_mode = "otsu"; // otsu | thres
    std::vector<double> thres;
    double jitter = 0; // fraction of edge length
    unsigned jitter_seed = 1234;
    int iters = 0; // number of segment() iterations before the dump
    std::string out = ".";
    int repeat = 3;
    int n_points_in_tet = 0;
};

static Args parse(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", k.c_str()); exit(2); }
            return argv[++i];
        };
        if (k == "--mode") a.mode = next();
        else if (k == "--volume") a.volume_path = next();
        else if (k == "--vol-dtype") a.vol_dtype = next();
        else if (k == "--dims") { a.dims[0] = atoi(next().c_str()); a.dims[1] = atoi(next().c_str()); a.dims[2] = atoi(next().c_str()); }
        else if (k == "--edge") a.edge = atof(next().c_str());
        else if (k == "--nb-phase") a.nb_phase = atoi(next().c_str());
        else if (k == "--alpha") a.alpha = atof(next().c_str());
        else if (k == "--labels") a.label_mode = next();
        else if (k == "--thres") { a.thres.clear(); std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) a.thres.push_back(atof(t.c_str())); }
        else if (k == "--jitter") a.jitter = atof(next().c_str());
        else if (k == "--jitter-seed") a.jitter_seed = (unsigned)atoi(next().c_str());
        else if (k == "--iters") a.iters = atoi(next().c_str());
        else if (k == "--out") a.out = next();
        else if (k == "--save-dsc") a.save_dsc = next();
        else if (k == "--repeat") a.repeat = atoi(next().c_str());
        else if (k == "--points-in-tet") a.n_points_in_tet = atoi(next().c_str());
        else { fprintf(stderr, "unknown arg %s\n", k.c_str()); exit(2); }
    }
    return a;
}

template <typename T>
static void write_bin(const std::string &dir, const std::string &name, const std::vector<T> &v)
{
    std::string p = dir + "/" + name;
    FILE *f = fopen(p.c_str(), "wb");
    if (!f) { perror(p.c_str()); exit(3); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

// Voxel decode rule shared with the product (include/dscgpu.h): u8 -> v/255.0 exactly as
// DEMO/image3d.cpp:127; u16 -> v/65535.0; f32 -> widened; f64 as is.
static void load_volume(const Args &a, image3d &img)
{
    size_t n = (size_t)a.dims[0] * a.dims[1] * a.dims[2];
    img._dim[0] = a.dims[0]; img._dim[1] = a.dims[1]; img._dim[2] = a.dims[2];
    img._voxels.resize(n);
    FILE *f = fopen(a.volume_path.c_str(), "rb");
    if (!f) { perror(a.volume_path.c_str()); exit(3); }
    if (a.vol_dtype == "u8") {
        std::vector<uint8_t> b(n); if (fread(b.data(), 1, n, f) != n) { fprintf(stderr, "short volume\n"); exit(3); }
        for (size_t i = 0; i < n; i++) img._voxels[i] = (double)b[i] / 255.0;
    } else if (a.vol_dtype == "u16") {
        std::vector<uint16_t> b(n); if (fread(b.data(), 2, n, f) != n) { fprintf(stderr, "short volume\n"); exit(3); }
        for (size_t i = 0; i < n; i++) img._voxels[i] = (double)b[i] / 65535.0;
    } else if (a.vol_dtype == "f32") {
        std::vector<float> b(n); if (fread(b.data(), 4, n, f) != n) { fprintf(stderr, "short volume\n"); exit(3); }
        for (size_t i = 0; i < n; i++) img._voxels[i] = (double)b[i];
    } else if (a.vol_dtype == "f64") {
        if (fread(img._voxels.data(), 8, n, f) != n) { fprintf(stderr, "short volume\n"); exit(3); }
    } else { fprintf(stderr, "bad dtype\n"); exit(2); }
    fclose(f);
    set_size();            // DEMO/tet_dis_coord.cpp:3129-3144 (image3d::load calls it at image3d.cpp:133)
    img.build_sum_table(); // DEMO/image3d.cpp:143-177
}

// Restatement of UI::init_dsc (DEMO/user_interface.cpp:550-630): regular grid, 6 tets per cube,
// padded by one edge length.  Optional deterministic jitter of nodes that are at least two grid
// layers away from the mesh boundary (the reference has no such option; it is used to obtain
// non-axis-aligned tets for the parity fixtures).
static std::unique_ptr<dsc_class> make_grid_dsc(const Args &a, const image3d &img)
{
    vec3 obj_dim = img.dimension_v();
    vec3 dsc_dim = obj_dim + vec3(2 * a.edge);
    double delta = a.edge;
    int NX = round(dsc_dim[0] / delta) + 1;
    int NY = round(dsc_dim[1] / delta) + 1;
    int NZ = round(dsc_dim[2] / delta) + 1;
    double deltax = dsc_dim[0] / (double)(NX - 1);
    double deltay = dsc_dim[1] / (double)(NY - 1);
    double deltaz = dsc_dim[2] / (double)(NZ - 1);

    std::vector<vec3> points;
    std::vector<int> tets;
    std::mt19937 rng(a.jitter_seed);
    std::uniform_real_distribution<double> uni(-1.0, 1.0);
    for (int iz = 0; iz < NZ; iz++)
        for (int iy = 0; iy < NY; iy++)
            for (int ix = 0; ix < NX; ix++) {
                vec3 p = vec3(ix * deltax, iy * deltay, iz * deltaz) - vec3(a.edge);
                if (a.jitter > 0) {
                    double jx = uni(rng), jy = uni(rng), jz = uni(rng);
                    bool interior = ix >= 2 && iy >= 2 && iz >= 2 && ix <= NX - 3 && iy <= NY - 3 && iz <= NZ - 3;
                    if (interior) p += vec3(jx, jy, jz) * (a.jitter * delta);
                }
                points.push_back(p);
            }
#define index_cube(x, y, z) ((z) * NX * NY + (y) * NX + (x))
    for (int iz = 0; iz < NZ - 1; iz++)
        for (int iy = 0; iy < NY - 1; iy++)
            for (int ix = 0; ix < NX - 1; ix++) {
                int vertices[] = {index_cube(ix, iy, iz), index_cube(ix + 1, iy, iz), index_cube(ix + 1, iy + 1, iz), index_cube(ix, iy + 1, iz),
                                  index_cube(ix, iy, iz + 1), index_cube(ix + 1, iy, iz + 1), index_cube(ix + 1, iy + 1, iz + 1), index_cube(ix, iy + 1, iz + 1)};
                int tetras[] = {0, 4, 5, 7, 0, 7, 5, 1, 0, 1, 3, 7, 1, 5, 6, 7, 1, 6, 7, 3, 1, 2, 6, 3};
                for (int i = 0; i < 24; i++) tets.push_back(vertices[tetras[i]]);
            }
    std::vector<int> tet_labels(tets.size() / 4, 0);
    std::unique_ptr<dsc_class> dsc(new dsc_class(points, tets, tet_labels));
    dsc->set_avg_edge_length(delta);
    return dsc;
}

// Restatement of UI::set_dsc_boundary_layer (DEMO/user_interface.cpp:441-462).
static void set_boundary_layer(dsc_class *dsc)
{
    std::vector<bool> is_tet_bound(dsc->get_no_tets_buffer(), false);
    for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++)
        if (nit->is_boundary())
            for (auto t : dsc->get_tets(nit.key())) is_tet_bound[t] = true;
    for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++)
        if (!is_tet_bound[tit.key()]) dsc->set_label(tit.key(), 1);
}

// Fixed-threshold labelling: same rule as segment_function.cpp:250-260 but with caller-supplied
// thresholds on the per-tet mean instead of Otsu's.
static void label_by_thresholds(segment_function &seg, const std::vector<double> &thres)
{
    dsc_class *dsc = seg._dsc;
    for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) {
        if (dsc->get_label(tit.key()) == BOUND_LABEL) continue;
        auto pos = dsc->get_pos(dsc->get_nodes(tit.key()));
        double mean = seg._img.get_tetra_intensity(pos);
        int label = 0;
        for (double t : thres) if (mean > t) label++;
        dsc->set_label(tit.key(), label + 1);
    }
}

struct CoutCapture {
    std::streambuf *old;
    std::ostringstream ss;
    CoutCapture() { old = std::cout.rdbuf(ss.rdbuf()); }
    ~CoutCapture() { std::cout.rdbuf(old); }
};

// hash of the voxel indices a tet's sample points fall into, computed with the reference's own
// table, macro and truncation (DEMO/image3d.cpp:351-367): order-sensitive FNV-1a over the
// (x,y,z) int triples, so a single mis-truncated coordinate changes it.
static uint64_t sample_hash(image3d &img, std::vector<vec3> tet_points, int *n_out)
{
    double v = Util::volume<double>(tet_points[0], tet_points[1], tet_points[2], tet_points[3]);
    long dis = std::upper_bound(dis_coord_size.begin(), dis_coord_size.end(), v) - dis_coord_size.begin() - 1;
    if (dis < 0) dis = 0;
    auto const &a = tet_dis_coord[dis];
    uint64_t h = 1469598103934665603ull;
    for (auto tb : a) {
        auto pt = get_coord(tet_points, tb);
        const int c[3] = {(int)pt[0], (int)pt[1], (int)pt[2]};
        for (int k = 0; k < 3; k++) { h ^= (uint64_t)(uint32_t)c[k]; h *= 1099511628211ull; }
    }
    *n_out = (int)a.size();
    return h;
}

static bool is_point_inside_ref(const vec3 &p, const vec3 &a, const vec3 &b, const vec3 &c, const vec3 &d)
{   // DEMO/segment_function.cpp:2576-2580
    auto bary_coord = Util::barycentric_coords<real>(p, a, b, c, d);
    return (bary_coord[0] >= 0 && bary_coord[1] >= 0 && bary_coord[2] >= 0 && bary_coord[3] >= 0);
}

static void dump_state(const Args &a, segment_function &seg, const std::string &dir)
{
    dsc_class *dsc = seg._dsc;
    std::string cmd = "mkdir -p '" + dir + "'";
    if (system(cmd.c_str()) != 0) exit(3);

    // --- run the reference hot path on the current mesh -------------------------------------
    std::string log;
    double energy_data = 0, energy_area = 0;
    {
        CoutCapture cap;
        std::cout.precision(17);
        seg.update_vertex_boundary();
        seg.update_average_intensity();
        seg.compute_external_force();
        seg.compute_internal_force_2();
        seg.compute_energy();
        log = cap.ss.str();
    }
    {
        size_t p = log.rfind("=== Energy; area = ");
        if (p != std::string::npos) sscanf(log.c_str() + p, "=== Energy; area = %lf - %lf", &energy_data, &energy_area);
    }

    // --- flatten the mesh through the public getters, in ascending key order ---------------
    unsigned nnb = dsc->get_no_nodes_buffer();
    std::vector<double> pos(3 * (size_t)nnb, 0.0);
    std::vector<uint8_t> node_valid(nnb, 0), vertex_bound(nnb, 0);
    for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++) {
        unsigned k = nit.key();
        vec3 p = dsc->get_pos(nit.key());
        pos[3 * k] = p[0]; pos[3 * k + 1] = p[1]; pos[3 * k + 2] = p[2];
        node_valid[k] = 1;
    }
    for (unsigned i = 0; i < nnb && i < seg.m_vertex_bound.size(); i++) vertex_bound[i] = seg.m_vertex_bound[i] ? 1 : 0;

    std::vector<int> tet_index(dsc->get_no_tets_buffer(), -1);
    std::vector<uint32_t> tet_key, tet_nodes;
    std::vector<int32_t> tet_label;
    for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) {
        tet_index[tit.key()] = (int)tet_key.size();
        tet_key.push_back(tit.key());
        auto nids = dsc->get_nodes(tit.key());
        for (int k = 0; k < 4; k++) tet_nodes.push_back(nids[k]);
        tet_label.push_back(dsc->get_label(tit.key()));
    }
    size_t nt = tet_key.size();

    std::vector<uint32_t> face_key, face_nodes, face_tets;
    std::vector<uint8_t> face_is_boundary;
    size_t n_faces_total = 0;
    for (auto fit = dsc->faces_begin(); fit != dsc->faces_end(); fit++) {
        n_faces_total++;
        if (!fit->is_interface()) continue;
        face_key.push_back(fit.key());
        auto nids = dsc->get_nodes(fit.key());
        for (int k = 0; k < 3; k++) face_nodes.push_back(nids[k]);
        auto tids = dsc->get_tets(fit.key());
        face_tets.push_back((uint32_t)tet_index[tids[0]]);
        face_tets.push_back(tids.size() > 1 ? (uint32_t)tet_index[tids[1]] : 0xFFFFFFFFu);
        face_is_boundary.push_back(fit->is_boundary() ? 1 : 0);
    }

    // --- per-tet reference quantities (A4/A5/A6 of SURVEY 8a) -------------------------------
    int P = seg.NB_PHASE;
    std::vector<double> tet_total(nt, -1.0), tet_vol(nt, -1.0), tet_mean(nt, -1.0);
    std::vector<double> tet_variation(nt * (size_t)P, -1.0), tet_energy(nt * (size_t)P, -1.0);
    std::vector<uint64_t> tet_hash(nt, 0);
    std::vector<int32_t> tet_nsamp(nt, 0);
    for (size_t i = 0; i < nt; i++) {
        is_mesh::TetrahedronKey tk(tet_key[i]);
        auto p = dsc->get_pos(dsc->get_nodes(tk));
        int ns;
        tet_hash[i] = sample_hash(seg._img, p, &ns);
        tet_nsamp[i] = ns;
        if (tet_label[i] == BOUND_LABEL) continue;
        double tot, vol;
        tet_mean[i] = seg._img.get_tetra_intensity(p, &tot, &vol);
        tet_total[i] = tot; tet_vol[i] = vol;
        for (int k = 0; k < P; k++) {
            tet_variation[i * P + k] = seg._img.get_variation(p, seg._mean_intensities[k]);
            tet_energy[i * P + k] = seg._img.get_energy(p, seg._mean_intensities[k]);
        }
    }
    // --- the initialisation pass (segment_function.cpp:165-262) on the current mesh: the 256-bin histogram of
    //     the per-tet means (labelled tets; the reference's loop also visits the -1 entries of its key-indexed buffer and writes
    //     bin -256 for them, out of bounds: there is nothing to match for those), the reference's own otsu_muti on it
    //     (DEMO/otsu_multi.h:105-129) and the labels the rule at :250-260 assigns
    std::vector<int> init_hist(256, 0);
    for (size_t i = 0; i < nt; i++) {
        if (tet_label[i] == BOUND_LABEL) continue;
        int idx = (int)(tet_mean[i] * 256);
        if (idx > 255) idx = 255;
        init_hist[idx]++;
    }
    std::vector<int> init_thres = otsu_muti(init_hist, P);
    std::vector<int32_t> init_labels(nt, 0);
    for (size_t i = 0; i < nt; i++) {
        if (tet_label[i] == BOUND_LABEL) continue;
        int m255 = (int)(tet_mean[i] * 255);
        if (m255 >= 255) m255 = 255;
        auto v_pos = std::lower_bound(init_thres.begin(), init_thres.end(), m255);
        int label = (int)(v_pos == init_thres.end()) ? init_thres.size() : int(v_pos - init_thres.begin());
        init_labels[i] = label + 1;
    }

    // relabel energies with the neighbour-area term (segment_function.cpp:618-643), labels unchanged
    std::vector<double> tet_relabel_energy(nt * (size_t)P, -1.0);
    for (size_t i = 0; i < nt; i++) {
        if (tet_label[i] == BOUND_LABEL) continue;
        is_mesh::TetrahedronKey tk(tet_key[i]);
        for (int k = 0; k < P; k++) tet_relabel_energy[i * P + k] = seg.get_energy_tetrahedron(tk, k + 1);
    }

    // --- per-tet face adjacency as get_energy_tetrahedron walks it (segment_function.cpp:629-640): the 4 faces in get_faces(t)
    //     order, each with its nodes in get_nodes(f) order (what DSC::area uses, src/DSC.h:3216-3220) and its co-boundary
    //     tets in get_tets(f) order, as indices into the dumped tet arrays
    std::vector<uint32_t> tet_face_nodes(nt * 12, 0), tet_face_tets(nt * 8, 0xFFFFFFFFu);
    for (size_t i = 0; i < nt; i++) {
        is_mesh::TetrahedronKey tk(tet_key[i]);
        int fi = 0;
        for (auto fid : dsc->get_faces(tk)) {
            if (fi >= 4) break;
            auto nids = dsc->get_nodes(fid);
            for (int k = 0; k < 3; k++) tet_face_nodes[i * 12 + fi * 3 + k] = nids[k];
            auto tids = dsc->get_tets(fid);
            tet_face_tets[i * 8 + fi * 2] = (uint32_t)tet_index[tids[0]];
            if (tids.size() > 1) tet_face_tets[i * 8 + fi * 2 + 1] = (uint32_t)tet_index[tids[1]];
            fi++;
        }
    }

    // --- voxel-in-tet predicate goldens (segment_function.cpp:2576-2580) --------------------
    std::vector<uint32_t> pit_tet;
    std::vector<double> pit_pts;
    std::vector<uint8_t> pit_inside;
    if (a.n_points_in_tet > 0 && nt > 0) {
        std::mt19937 rng(777);
        std::uniform_real_distribution<double> uni(0.0, 1.0);
        for (int i = 0; i < a.n_points_in_tet; i++) {
            size_t t = (size_t)(uni(rng) * nt) % nt;
            is_mesh::TetrahedronKey tk(tet_key[t]);
            auto p = dsc->get_pos(dsc->get_nodes(tk));
            vec3 q;
            int kind = i % 4;
            if (kind == 0) { // voxel centre near the tet
                vec3 c = (p[0] + p[1] + p[2] + p[3]) * 0.25;
                q = vec3(std::floor(c[0] + (uni(rng) - 0.5) * a.edge) + 0.5, std::floor(c[1] + (uni(rng) - 0.5) * a.edge) + 0.5,
                         std::floor(c[2] + (uni(rng) - 0.5) * a.edge) + 0.5);
            } else if (kind == 1) { // exactly a vertex
                q = p[i % 4];
            } else if (kind == 2) { // on a face (one barycentric coordinate exactly zero before rounding)
                double u = uni(rng), v = uni(rng); if (u + v > 1) { u = 1 - u; v = 1 - v; }
                q = p[0] * u + p[1] * v + p[2] * (1 - u - v);
            } else { // random point in the bounding box
                vec3 lo = p[0], hi = p[0];
                for (int k = 1; k < 4; k++) for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], p[k][d]); hi[d] = std::max(hi[d], p[k][d]); }
                q = vec3(lo[0] + uni(rng) * (hi[0] - lo[0]), lo[1] + uni(rng) * (hi[1] - lo[1]), lo[2] + uni(rng) * (hi[2] - lo[2]));
            }
            pit_tet.push_back((uint32_t)t);
            pit_pts.push_back(q[0]); pit_pts.push_back(q[1]); pit_pts.push_back(q[2]);
            pit_inside.push_back(is_point_inside_ref(q, p[0], p[1], p[2], p[3]) ? 1 : 0);
        }
    }

    // --- outputs of the reference members ---------------------------------------------------
    std::vector<double> forces(3 * (size_t)nnb, 0.0);
    for (size_t i = 0; i < seg._forces.size() && i < nnb; i++) { forces[3 * i] = seg._forces[i][0]; forces[3 * i + 1] = seg._forces[i][1]; forces[3 * i + 2] = seg._forces[i][2]; }

    // compute_internal_force_2 and the node attributes it reads (is_mesh/attributes.h:84-102)
    std::vector<double> internal_forces(3 * (size_t)nnb, 0.0);
    for (size_t i = 0; i < seg._internal_forces.size() && i < nnb; i++)
        for (int d = 0; d < 3; d++) internal_forces[3 * i + d] = seg._internal_forces[i][d];
    std::vector<uint8_t> node_flags(nnb, 0);
    for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++)
        node_flags[(unsigned)nit.key()] = (uint8_t)((nit->is_interface() ? 1 : 0) | (nit->is_crossing() ? 2 : 0) | (nit->is_boundary() ? 4 : 0));
    write_bin(dir, "init_hist.i32", init_hist);
    write_bin(dir, "init_thres.i32", init_thres);
    write_bin(dir, "init_labels.i32", init_labels);
    write_bin(dir, "internal_forces.f64", internal_forces);
    write_bin(dir, "node_flags.u8", node_flags);
    // the order in which DSC::get_normal(node) (src/DSC.h:3138-3159) adds the face normals: the interface faces of get_faces(node)
    // (is_mesh/is_mesh.h:691-699; the DSC_CACHE copy is filled by the same walk, src/DSC.h:243-251), as indices into the interface-face list
    {
        std::vector<int32_t> face_index(dsc->get_no_faces_buffer(), -1);
        for (size_t i = 0; i < face_key.size(); i++) face_index[face_key[i]] = (int32_t)i;
        std::vector<uint32_t> order_off(nnb + 1, 0), order_idx;
        for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++) {
            const unsigned k = (unsigned)nit.key();
            if (nit->is_interface() && !nit->is_crossing())
                for (auto f : dsc->get_faces(nit.key()))
                    if (dsc->get(f).is_interface()) { order_idx.push_back((uint32_t)face_index[(unsigned)f]); order_off[k + 1]++; }
        }
        for (unsigned k = 0; k < nnb; k++) order_off[k + 1] += order_off[k];
        write_bin(dir, "node_face_off.u32", order_off);
        write_bin(dir, "node_face_idx.u32", order_idx);
    }

    write_bin(dir, "pos.f64", pos);
    write_bin(dir, "node_valid.u8", node_valid);
    write_bin(dir, "vertex_bound.u8", vertex_bound);
    write_bin(dir, "tet_key.u32", tet_key);
    write_bin(dir, "tet_nodes.u32", tet_nodes);
    write_bin(dir, "tet_label.i32", tet_label);
    write_bin(dir, "face_key.u32", face_key);
    write_bin(dir, "face_nodes.u32", face_nodes);
    write_bin(dir, "face_tets.u32", face_tets);
    write_bin(dir, "face_is_boundary.u8", face_is_boundary);
    write_bin(dir, "tet_total.f64", tet_total);
    write_bin(dir, "tet_vol.f64", tet_vol);
    write_bin(dir, "tet_mean.f64", tet_mean);
    write_bin(dir, "tet_variation.f64", tet_variation);
    write_bin(dir, "tet_energy.f64", tet_energy);
    write_bin(dir, "tet_relabel_energy.f64", tet_relabel_energy);
    write_bin(dir, "tet_face_nodes.u32", tet_face_nodes);
    write_bin(dir, "tet_face_tets.u32", tet_face_tets);
    write_bin(dir, "tet_hash.u64", tet_hash);
    write_bin(dir, "tet_nsamp.i32", tet_nsamp);
    write_bin(dir, "mean.f64", seg._mean_intensities);
    write_bin(dir, "total.f64", seg._total_intensities);
    write_bin(dir, "phase_volume.f64", seg._phase_volume);
    write_bin(dir, "forces.f64", forces);
    write_bin(dir, "pit_tet.u32", pit_tet);
    write_bin(dir, "pit_pts.f64", pit_pts);
    write_bin(dir, "pit_inside.u8", pit_inside);

    FILE *f = fopen((dir + "/meta.json").c_str(), "w");
    fprintf(f, "{\"dims\": [%d, %d, %d], \"vol_dtype\": \"%s\", \"edge\": %.17g, \"nb_phase\": %d, \"alpha\": %.17g,\n"
               " \"n_nodes_buffer\": %u, \"n_nodes\": %u, \"n_tets\": %zu, \"n_faces_total\": %zu, \"n_if\": %zu,\n"
               " \"energy_data\": %.17g, \"energy_area\": %.17g, \"iters\": %d, \"jitter\": %.17g}\n",
            seg._img._dim[0], seg._img._dim[1], seg._img._dim[2], a.vol_dtype.c_str(), a.edge, seg.NB_PHASE, seg.m_alpha,
            nnb, dsc->get_no_nodes(), nt, n_faces_total, face_key.size(), energy_data, energy_area, a.iters, a.jitter);
    fclose(f);
    fprintf(stderr, "[ref_driver] dumped %s: %u nodes(buffer) %zu tets %zu interface faces; mean:", dir.c_str(), nnb, nt, face_key.size());
    for (double m : seg._mean_intensities) fprintf(stderr, " %.12g", m);
    fprintf(stderr, "\n");
}

int main(int argc, char **argv)
{
    Args a = parse(argc, argv);
    segment_function seg;
    seg.NB_PHASE = a.nb_phase;
    seg.m_alpha = a.alpha;
    seg.VARIATION_THRES_FOR_RELABELING = 0.5;
    load_volume(a, seg._img);

    double t0 = now_s();
    std::unique_ptr<dsc_class> dsc = make_grid_dsc(a, seg._img);
    double t_mesh = now_s() - t0;
    dsc->set_min_edge_length(a.edge); // DEMO/user_interface.cpp:272 (DSC::set_min_edge_length via UI)
    seg.set_min_edge_length(a.edge);
    set_boundary_layer(dsc.get());
    seg._dsc = dsc.get();
    {
        CoutCapture cap;
        if (a.label_mode == "otsu") seg.initialization_discrete_opt();
        else label_by_thresholds(seg, a.thres);
    }

    if (a.mode == "dump") {
        for (int it = 0; it < a.iters; it++) {
            CoutCapture cap;
            seg.segment();
        }
        dump_state(a, seg, a.out);
        if (!a.save_dsc.empty()) { // what DSC "save" does: extract_tet_mesh (is_mesh/is_mesh.h:1686-1706) + export_tet_mesh
            std::vector<vec3> points;
            std::vector<int> tets, tet_labels;
            dsc->extract_tet_mesh(points, tets, tet_labels);
            is_mesh::export_tet_mesh(a.save_dsc, points, tets, tet_labels);
        }
        return 0;
    }
    if (a.mode == "time") {
        // wall-clock of the reference members on this host, single thread (the reference is
        // single-threaded on this path: SURVEY.md section 2)
        double best_avg = 1e30, best_force = 1e30, best_tet = 1e30, best_energy = 1e30;
        size_t n_tets_lab = 0, n_if = 0, nt = 0;
        for (auto fit = dsc->faces_begin(); fit != dsc->faces_end(); fit++) if (fit->is_interface()) n_if++;
        {
            CoutCapture cap;
            seg.update_vertex_boundary();
            for (int r = 0; r < a.repeat; r++) {
                double s = now_s(); seg.update_average_intensity(); best_avg = std::min(best_avg, now_s() - s);
                s = now_s(); seg.compute_external_force(); best_force = std::min(best_force, now_s() - s);
                s = now_s();
                n_tets_lab = 0; nt = 0;
                double acc = 0;
                std::vector<double> tot(a.nb_phase, 0.0), vol(a.nb_phase, 0.0);
                for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) {
                    nt++;
                    int l = dsc->get_label(tit.key());
                    if (l == BOUND_LABEL) continue;
                    double t, v;
                    acc += seg._img.get_tetra_intensity(dsc->get_pos(dsc->get_nodes(tit.key())), &t, &v);
                    tot[l - 1] += t; vol[l - 1] += v;
                    n_tets_lab++;
                }
                best_tet = std::min(best_tet, now_s() - s);
                if (acc < -1) printf("%f", acc);
                s = now_s(); seg.compute_energy(); best_energy = std::min(best_energy, now_s() - s);
            }
        }
        printf("{\"n_tets\": %zu, \"n_tets_labelled\": %zu, \"n_if\": %zu, \"t_mesh_build_s\": %.6f, \"t_zray_means_s\": %.6f, "
               "\"t_external_force_s\": %.6f, \"t_tet_pass_s\": %.6f, \"t_energy_s\": %.6f, \"repeat\": %d}\n",
               nt, n_tets_lab, n_if, t_mesh, best_avg, best_force, best_tet, best_energy, a.repeat);
        return 0;
    }
    fprintf(stderr, "unknown mode\n");
    return 2;
}
