// oracle/dropin_test.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The drop-in demonstration: the UNMODIFIED reference sources (same objects as oracle/_ref/dsc_ref) and the product's
// C ABI (libdscgpu.so, through include/dscgpu_adapter.hpp) in ONE process, on the same live is_mesh/DSC complex.
// For the initial mesh and after every segment() iteration (deform + relabel + adapt make the mesh irregular) it runs
//     reference: update_vertex_boundary(); update_average_intensity(); compute_external_force(); compute_internal_force_2(); compute_energy()
//     GPU      : adapter.refresh(); adapter.update_average_intensity(); adapter.compute_external_force(); compute_energy()
// on the same segment_function object and compares the public result members:
//     _phase_volume bit-exact, _total_intensities/_mean_intensities 1e-6 relative, _forces bit-exact when fed the
//     reference's means (gather kernel keeps the reference's addition order) and 1e-6 with its own means, energy 1e-6.
// Exit code 0 = parity.  Needs a CUDA device (run under tests/test_dropin_gpu.py on the GPU box).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <memory>
#include <iostream>
#include <sstream>

#include "segment_function.h"
#include "tet_dis_coord.hpp"
#include "dscgpu_adapter.hpp"

// DEMO/otsu_multi.h defines (not declares) its functions; the definition comes from the reference's segment_function.o
std::vector<int> otsu_muti(std::vector<int> histogram_in, int classes);

double noise_level = 0;
int num_images = 0;
bool arg_b_build_table_origin = true;

typedef DSC::DeformableSimplicialComplex<> dsc_class;

struct Quiet {
    std::streambuf *old;
    std::ostringstream ss;
    Quiet() { old = std::cout.rdbuf(ss.rdbuf()); }
    ~Quiet() { std::cout.rdbuf(old); }
};

static std::unique_ptr<dsc_class> make_grid(const image3d &img, double edge)
{   // UI::init_dsc, DEMO/user_interface.cpp:550-630
    vec3 dsc_dim = img.dimension_v() + vec3(2 * edge);
    int NX = round(dsc_dim[0] / edge) + 1, NY = round(dsc_dim[1] / edge) + 1, NZ = round(dsc_dim[2] / edge) + 1;
    double dx = dsc_dim[0] / (double)(NX - 1), dy = dsc_dim[1] / (double)(NY - 1), dz = dsc_dim[2] / (double)(NZ - 1);
    std::vector<vec3> points;
    std::vector<int> tets;
    for (int iz = 0; iz < NZ; iz++)
        for (int iy = 0; iy < NY; iy++)
            for (int ix = 0; ix < NX; ix++) points.push_back(vec3(ix * dx, iy * dy, iz * dz) - vec3(edge));
#define IC(x, y, z) ((z) * NX * NY + (y) * NX + (x))
    for (int iz = 0; iz < NZ - 1; iz++)
        for (int iy = 0; iy < NY - 1; iy++)
            for (int ix = 0; ix < NX - 1; ix++) {
                int v[] = {IC(ix, iy, iz), IC(ix + 1, iy, iz), IC(ix + 1, iy + 1, iz), IC(ix, iy + 1, iz),
                           IC(ix, iy, iz + 1), IC(ix + 1, iy, iz + 1), IC(ix + 1, iy + 1, iz + 1), IC(ix, iy + 1, iz + 1)};
                int tt[] = {0, 4, 5, 7, 0, 7, 5, 1, 0, 1, 3, 7, 1, 5, 6, 7, 1, 6, 7, 3, 1, 2, 6, 3};
                for (int i = 0; i < 24; i++) tets.push_back(v[tt[i]]);
            }
    std::vector<int> labels(tets.size() / 4, 0);
    std::unique_ptr<dsc_class> dsc(new dsc_class(points, tets, labels));
    dsc->set_avg_edge_length(edge);
    return dsc;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool close_rel(double a, double b, double tol) { return std::fabs(a - b) <= tol * std::max(std::fabs(a), std::fabs(b)) + 1e-300; }

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: dsc_dropin volume.u8 X Y Z edge iters [nb_phase alpha]\n"); return 2; }
    const int X = atoi(argv[2]), Y = atoi(argv[3]), Z = atoi(argv[4]);
    const double edge = atof(argv[5]);
    const int iters = atoi(argv[6]);
    segment_function seg;
    seg.NB_PHASE = argc > 7 ? atoi(argv[7]) : 3;
    seg.m_alpha = argc > 8 ? atof(argv[8]) : 0.001;
    seg.VARIATION_THRES_FOR_RELABELING = 0.5;
    {
        size_t n = (size_t)X * Y * Z;
        std::vector<uint8_t> b(n);
        FILE *f = fopen(argv[1], "rb");
        if (!f || fread(b.data(), 1, n, f) != n) { fprintf(stderr, "cannot read volume\n"); return 2; }
        fclose(f);
        seg._img._dim[0] = X; seg._img._dim[1] = Y; seg._img._dim[2] = Z;
        seg._img._voxels.resize(n);
        for (size_t i = 0; i < n; i++) seg._img._voxels[i] = (double)b[i] / 255.0;
        set_size();
        seg._img.build_sum_table();
    }
    std::unique_ptr<dsc_class> dsc = make_grid(seg._img, edge);
    dsc->set_min_edge_length(edge);
    seg.set_min_edge_length(edge);
    {   // UI::set_dsc_boundary_layer, DEMO/user_interface.cpp:441-462
        std::vector<bool> bound(dsc->get_no_tets_buffer(), false);
        for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++)
            if (nit->is_boundary()) for (auto t : dsc->get_tets(nit.key())) bound[t] = true;
        for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) if (!bound[tit.key()]) dsc->set_label(tit.key(), 1);
    }
    seg._dsc = dsc.get();
    // The reference's pass runs first and before anything else is allocated: its histogram loop also visits the -1 entries of
    // its key-indexed buffer and increments bin -256 (DEMO/segment_function.cpp:219-226), 1 KB in front of the vector -- a
    // latent out-of-bounds write (AddressSanitizer flags it) whose victim depends on what the heap holds at that moment.
    { Quiet q; seg.initialization_discrete_opt(); }
    std::vector<int> lab_ref;
    for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) lab_ref.push_back(dsc->get_label(tit.key()));

    int failures = 0;
    try {
        dscgpu::ImageEnergy<image3d, segment_function> gpu(seg._img, 0);
        {   // the same pass through the GPU, on the mesh as it was before (boundary layer 0, everything else 1)
            for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++)
                if (dsc->get_label(tit.key()) != 0) dsc->set_label(tit.key(), 1);
            gpu.refresh(seg);
            gpu.initialization_discrete_opt(seg, otsu_muti);
            size_t i = 0, bad = 0, n = 0;
            for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++, i++) {
                n++;
                if (lab_ref[i] != dsc->get_label(tit.key())) { bad++; dsc->set_label(tit.key(), lab_ref[i]); }
            }
            // the per-tet means of this pass are added in the reference's order: every label must be the reference's
            const bool ok = bad == 0;
            printf("initialisation pass: %zu tets, %zu labels differ from the reference's | %s\n", n, bad, ok ? "OK" : "FAIL");
            if (!ok) failures++;
        }
        for (int it = 0; it <= iters; it++) {
            // ---- reference ----
            double t_ref[6] = {0, 0, 0, 0, 0, 0}, t_gpu[6] = {0, 0, 0, 0, 0, 0};
            std::string log;
            {
                Quiet q;
                std::cout.precision(17);
                double t = now_ms();
                seg.update_vertex_boundary();       t_ref[0] = now_ms() - t; t = now_ms();
                seg.update_average_intensity();     t_ref[1] = now_ms() - t; t = now_ms();
                seg.compute_external_force();       t_ref[2] = now_ms() - t; t = now_ms();
                seg.compute_internal_force_2();     t_ref[3] = now_ms() - t; t = now_ms();
                seg.compute_energy();               t_ref[4] = now_ms() - t;
                log = q.ss.str();
            }
            double e_data_ref = 0, e_area_ref = 0;
            size_t p = log.rfind("=== Energy; area = ");
            if (p != std::string::npos) sscanf(log.c_str() + p, "=== Energy; area = %lf - %lf", &e_data_ref, &e_area_ref);
            const std::vector<double> mean_ref = seg._mean_intensities, total_ref = seg._total_intensities, vol_ref = seg._phase_volume;
            const std::vector<vec3> forces_ref = seg._forces, internal_ref = seg._internal_forces;
            // ---- GPU, same object, same mesh ----
            { double t = now_ms(); gpu.refresh(seg); t_gpu[0] = now_ms() - t; }
            {   // m_vertex_bound from the device against the reference's (update_vertex_boundary ran above)
                const std::vector<bool> vb_ref = seg.m_vertex_bound;
                gpu.update_vertex_boundary_flags(seg);
                size_t bad_vb = vb_ref.size() != seg.m_vertex_bound.size() ? 1 : 0;
                for (size_t i = 0; i < vb_ref.size() && !bad_vb; i++) if (vb_ref[i] != seg.m_vertex_bound[i]) bad_vb++;
                printf("iter %d: m_vertex_bound %s\n", it, bad_vb ? "DIFFERENT" : "flags identical");
                if (bad_vb) failures++;
                seg.m_vertex_bound = vb_ref;
            }
            { double t = now_ms(); gpu.compute_internal_force_2(seg); t_gpu[3] = now_ms() - t; } // SURVEY 8(f) row 3 with the get_faces(node) order of the complex: bit for bit
            double imax = 0, idmax = 0;
            size_t bad_i = seg._internal_forces.size() != internal_ref.size() ? 1 : 0;
            for (size_t i = 0; i < internal_ref.size() && !bad_i; i++) for (int c = 0; c < 3; c++) {
                imax = std::max(imax, std::fabs(internal_ref[i][c]));
                idmax = std::max(idmax, std::fabs(internal_ref[i][c] - seg._internal_forces[i][c]));
            }
            seg._internal_forces = internal_ref;
            { double t = now_ms(); gpu.compute_external_force(seg); t_gpu[2] = now_ms() - t; } // with the reference's means still in seg._mean_intensities: bit-exact bar
            size_t bad_f = 0;
            if (seg._forces.size() != forces_ref.size()) bad_f = 1;
            else for (size_t i = 0; i < forces_ref.size(); i++) for (int c = 0; c < 3; c++) if (seg._forces[i][c] != forces_ref[i][c]) bad_f++;
            { double t = now_ms(); gpu.update_average_intensity(seg); t_gpu[1] = now_ms() - t; }
            bool ok = bad_f == 0;
            for (int k = 0; k < seg.NB_PHASE; k++) {
                if (seg._phase_volume[k] != vol_ref[k]) ok = false;
                if (!close_rel(seg._total_intensities[k], total_ref[k], 1e-6)) ok = false;
                if (!close_rel(seg._mean_intensities[k], mean_ref[k], 1e-6)) ok = false;
            }
            gpu.compute_external_force(seg); // with the GPU's own means: tolerance bar
            double fmax = 0, dmax = 0;
            for (size_t i = 0; i < forces_ref.size(); i++) for (int c = 0; c < 3; c++) {
                fmax = std::max(fmax, std::fabs(forces_ref[i][c]));
                dmax = std::max(dmax, std::fabs(forces_ref[i][c] - seg._forces[i][c]));
            }
            if (dmax > 1e-6 * fmax) ok = false;
            double e_data = 0, e_area = 0;
            { double t = now_ms(); gpu.compute_energy(seg, &e_data, &e_area); t_gpu[4] = now_ms() - t; }
            if (!close_rel(e_data, e_data_ref, 1e-6) || !close_rel(e_area, e_area_ref, 1e-6)) ok = false;
            if (bad_i || idmax != 0.0) ok = false;
            printf("iter %d: internal forces max abs diff %.3g of %.3g (get_faces(node) walk %.2f ms of the %.2f ms call)\n", it, idmax, imax, gpu.last_face_order_ms(), t_gpu[3]);
            printf("iter %d: %zu tets %zu interface faces | phase_volume %s | forces(ref means) %zu mismatching components | "
                   "forces(gpu means) max abs diff %.3g of %.3g | energy %.10g/%.10g vs %.10g/%.10g | %s\n",
                   it, gpu.snapshot().tet_label.size(), gpu.snapshot().face_is_boundary.size(),
                   (seg._phase_volume == vol_ref) ? "bit-exact" : "DIFFERENT", bad_f, dmax, fmax, e_data, e_area, e_data_ref, e_area_ref,
                   ok ? "OK" : "FAIL");
            if (!ok) failures++;
            printf("iter %d: snapshot sent %s", it, gpu.last_refresh_incremental() ? "incrementally" : "in full");
            if (gpu.last_refresh_incremental()) printf(" (%zu moved nodes, %zu changed tets)", gpu.last_moved_nodes(), gpu.last_changed_tets());
            printf("\n");
            // restore the reference's results
            seg._mean_intensities = mean_ref; seg._total_intensities = total_ref; seg._phase_volume = vol_ref; seg._forces = forces_ref;
            {   // SURVEY 8(f) row 3, second half: get_node_displacement + the non-topological part of snapp_boundary.  The reference
                // first (it ends in set_destination for every interface node), then the device with the same forces.
                const size_t nnb = dsc->get_no_nodes_buffer();
                const double dt0 = seg._dt;
                // per-node state the call reads and then rewrites (adaptive time step, :545-578): both sides must start from the same
                seg._dt_adapt.resize(nnb, 1.0); seg._previous_dis.resize(nnb, vec3(0)); seg._cur_dis.resize(nnb, vec3(0));
                const std::vector<double> adapt0 = seg._dt_adapt;
                const std::vector<vec3> prev0 = seg._previous_dis, cur0 = seg._cur_dis;
                std::vector<char> clean0(nnb, 0);
                for (size_t i = 0; i < nnb; i++) clean0[i] = dsc->cache.is_clean[i] != nullptr;
                auto restore_clean = [&]() {
                    for (size_t i = 0; i < nnb; i++)
                        if (!clean0[i] && dsc->cache.is_clean[i]) { delete dsc->cache.is_clean[i]; dsc->cache.is_clean[i] = nullptr; }
                };
                double t = now_ms();
                { Quiet q; seg.snapp_boundary(); }
                t_ref[5] = now_ms() - t;
                const std::vector<double> adapt_ref = seg._dt_adapt;
                const double dt_ref = seg._dt;
                std::vector<vec3> dest_ref(nnb, vec3(0));
                for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++) dest_ref[(unsigned)nit.key()] = dsc->get(nit.key()).get_destination();
                seg._dt = dt0;
                // (a) the kernels on the reference's own force arrays: bit for bit
                std::vector<double> F(3 * nnb), Fi(3 * nnb), dest(3 * nnb), cur(3 * nnb), adapt(adapt0);
                std::vector<uint8_t> vb(nnb, 0);
                for (size_t i = 0; i < nnb; i++) for (int c = 0; c < 3; c++) { F[3 * i + c] = forces_ref[i][c]; Fi[3 * i + c] = internal_ref[i][c]; }
                for (size_t i = 0; i < nnb && i < seg.m_vertex_bound.size(); i++) vb[i] = seg.m_vertex_bound[i] ? 1 : 0;
                static const double thr = dsc->AVG_LENGTH * 0.5 * 0.6, maxd = dsc->get_avg_edge_length() * 0.5 * 0.4;
                dscgpu_move_params mp; mp.dt = dt0; mp.alpha = seg.m_alpha; mp.align_threshold = thr; mp.max_displacement = maxd;
                dscgpu_move_stats st;
                if (dscgpu_node_destinations(gpu.ctx(), &mp, F.data(), Fi.data(), gpu.snapshot().node_flags.data(), vb.data(), adapt.data(), dest.data(), cur.data(), &st) != DSCGPU_OK)
                    throw std::runtime_error(dscgpu_last_error(gpu.ctx()));
                size_t bad_d = 0, n_if = 0;
                for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++) {
                    if (!nit->is_interface()) continue;
                    n_if++;
                    const unsigned k = nit.key();
                    dsc->set_destination(nit.key(), vec3(dest[3 * k], dest[3 * k + 1], dest[3 * k + 2])); // the same clamping the reference's value went through
                    const vec3 d = dsc->get(nit.key()).get_destination();
                    for (int c = 0; c < 3; c++) if (d[c] != dest_ref[k][c]) bad_d++;
                }
                const bool dt_ok = dt0 * st.scale == dt_ref;
                // (b) the adapter's body on the forces resident on the device (internal forces of projected nodes agree to rounding only)
                gpu.compute_external_force(seg);
                gpu.compute_internal_force_2(seg);
                seg._dt_adapt = adapt0; seg._previous_dis = prev0; seg._cur_dis = cur0; restore_clean();
                t = now_ms();
                gpu.snapp_boundary(seg);
                t_gpu[5] = now_ms() - t;
                size_t bad_adapt = 0;
                for (size_t i = 0; i < nnb; i++) if (seg._dt_adapt[i] != adapt_ref[i]) bad_adapt++;
                double ddmax = 0;
                for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++) {
                    if (!nit->is_interface()) continue;
                    const vec3 d = dsc->get(nit.key()).get_destination();
                    for (int c = 0; c < 3; c++) ddmax = std::max(ddmax, std::fabs(d[c] - dest_ref[(unsigned)nit.key()][c]));
                }
                seg._dt = dt0; seg._forces = forces_ref; seg._internal_forces = internal_ref;
                // (a step factor _dt_adapt may flip when a direction cosine sits within rounding of its threshold: counted, not required to be 0)
                const bool ok_d = bad_d == 0 && dt_ok && ddmax < 1e-9 && bad_adapt * 100 <= n_if;
                printf("iter %d: destinations of %zu interface nodes: %zu mismatching components (reference forces) %s | adapter body max abs diff %.3g, "
                       "%zu step factors differ | dt scale %s | %s\n",
                       it, n_if, bad_d, bad_d == 0 ? "destinations bit-exact" : "DIFFERENT", ddmax, bad_adapt, dt_ok ? "equal" : "DIFFERENT", ok_d ? "OK" : "FAIL");
                if (!ok_d) failures++;
                // the state segment() would find had it been alone
                seg._dt_adapt = adapt0; seg._previous_dis = prev0; seg._cur_dis = cur0; seg._dt = dt0; restore_clean();
            }
            printf("iter %d timing [ms] reference: update_vertex_boundary %.2f update_average_intensity %.2f compute_external_force %.2f compute_internal_force_2 %.2f "
                   "compute_energy %.2f snapp_boundary %.2f | gpu drop-in: refresh %.2f (is_mesh walk %.2f, %s) update_average_intensity %.2f compute_external_force %.2f "
                   "compute_internal_force_2 %.2f compute_energy %.2f snapp_boundary %.2f\n",
                   it, t_ref[0], t_ref[1], t_ref[2], t_ref[3], t_ref[4], t_ref[5], t_gpu[0], gpu.last_flatten_ms(), gpu.last_refresh_incremental() ? "incremental" : "full",
                   t_gpu[1], t_gpu[2], t_gpu[3], t_gpu[4], t_gpu[5]);
            // let the reference move the mesh
            if (it < iters) { Quiet q; seg.segment(); }
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "drop-in test failed: %s\n", e.what());
        return 3;
    }
    printf("%s\n", failures ? "DROP-IN PARITY FAILED" : "DROP-IN PARITY OK");
    return failures ? 1 : 0;
}
