// oracle/walk_timing.cpp -- TEST INFRASTRUCTURE ONLY (CPU, no CUDA device needed).
//
// Times the host side of the drop-in on a live is_mesh/DSC complex built from the UNMODIFIED reference sources: the walk that flattens
// the complex into the snapshot arrays (dscgpu::snapshot_from_dsc, include/dscgpu_adapter.hpp), split into its node, tet and face parts.
//   usage: dsc_walk volume.u8 X Y Z edge [repeats]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <memory>
#include <iostream>
#include <sstream>
#include <vector>

#include "segment_function.h"
#include "tet_dis_coord.hpp"
#include "dscgpu_adapter.hpp"

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
{   // the grid of UI::init_dsc (DEMO/user_interface.cpp:550-630): six tets per cube
    vec3 dsc_dim = img.dimension_v() + vec3(2 * edge);
    int NX = round(dsc_dim[0] / edge) + 1, NY = round(dsc_dim[1] / edge) + 1, NZ = round(dsc_dim[2] / edge) + 1;
    double dx = dsc_dim[0] / (double)(NX - 1), dy = dsc_dim[1] / (double)(NY - 1), dz = dsc_dim[2] / (double)(NZ - 1);
    std::vector<vec3> points;
    std::vector<int> tets;
    for (int iz = 0; iz < NZ; iz++)
        for (int iy = 0; iy < NY; iy++)
            for (int ix = 0; ix < NX; ix++) points.push_back(vec3(ix * dx, iy * dy, iz * dz) - vec3(edge));
    auto ic = [&](int x, int y, int z) { return z * NX * NY + y * NX + x; };
    for (int iz = 0; iz < NZ - 1; iz++)
        for (int iy = 0; iy < NY - 1; iy++)
            for (int ix = 0; ix < NX - 1; ix++) {
                int v[] = {ic(ix, iy, iz), ic(ix + 1, iy, iz), ic(ix + 1, iy + 1, iz), ic(ix, iy + 1, iz),
                           ic(ix, iy, iz + 1), ic(ix + 1, iy, iz + 1), ic(ix + 1, iy + 1, iz + 1), ic(ix, iy + 1, iz + 1)};
                int tt[] = {0, 4, 5, 7, 0, 7, 5, 1, 0, 1, 3, 7, 1, 5, 6, 7, 1, 6, 7, 3, 1, 2, 6, 3};
                for (int i = 0; i < 24; i++) tets.push_back(v[tt[i]]);
            }
    std::vector<int> labels(tets.size() / 4, 0);
    std::unique_ptr<dsc_class> dsc(new dsc_class(points, tets, labels));
    dsc->set_avg_edge_length(edge);
    return dsc;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: dsc_walk volume.u8 X Y Z edge [repeats]\n"); return 2; }
    const int X = atoi(argv[2]), Y = atoi(argv[3]), Z = atoi(argv[4]);
    const double edge = atof(argv[5]);
    const int reps = argc > 6 ? atoi(argv[6]) : 3;
    segment_function seg;
    seg.NB_PHASE = 3;
    seg.m_alpha = 0.001;
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
    {
        std::vector<bool> bound(dsc->get_no_tets_buffer(), false);
        for (auto nit = dsc->nodes_begin(); nit != dsc->nodes_end(); nit++)
            if (nit->is_boundary()) for (auto t : dsc->get_tets(nit.key())) bound[t] = true;
        for (auto tit = dsc->tetrahedra_begin(); tit != dsc->tetrahedra_end(); tit++) if (!bound[tit.key()]) dsc->set_label(tit.key(), 1);
    }
    seg._dsc = dsc.get();
    { Quiet q; seg.initialization_discrete_opt(); }
    for (int r = 0; r < reps; r++) {
        dscgpu::Snapshot s;
        dscgpu::WalkTimes wt;
        const double t0 = now_ms();
        dscgpu::snapshot_from_dsc(*dsc, s, /*key_indexed=*/true, &wt);
        const double t = now_ms() - t0;
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void *p, size_t n) { const unsigned char *c = static_cast<const unsigned char *>(p); for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; } };
        mix(s.pos.data(), s.pos.size() * 8); mix(s.tet_nodes.data(), s.tet_nodes.size() * 4); mix(s.tet_label.data(), s.tet_label.size() * 4);
        mix(s.face_nodes.data(), s.face_nodes.size() * 4); mix(s.face_tets.data(), s.face_tets.size() * 4); mix(s.face_is_boundary.data(), s.face_is_boundary.size());
        mix(s.node_flags.data(), s.node_flags.size()); mix(s.face_key.data(), s.face_key.size() * 4);
        printf("walk %d: %.2f ms (nodes %.2f, tets %.2f, faces %.2f) | %zu tet slots, %zu interface faces, %zu nodes | snapshot hash %016llx\n", r, t, wt.nodes_ms, wt.tets_ms,
               wt.faces_ms, s.tet_label.size(), s.face_is_boundary.size(), s.pos.size() / 3, (unsigned long long)h);
    }
    return 0;
}
