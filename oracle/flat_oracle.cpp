// oracle/flat_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Flat CPU restatement ("oracle B" of SURVEY.md 8c) of the reference's image-energy path over
// the SoA mesh snapshot that the product's C ABI consumes (include/dscgpu.h).  It exists so
// that (1) parity tests have an independent checker at sizes where building an is_mesh
// complex is impractical and on the GPU box where /root/reference does not exist, and
// (2) bench.py has a multi-threaded CPU baseline.  It is pinned against the unmodified
// reference (oracle/_ref/dsc_ref, built by oracle/Makefile) through the fixtures in
// tests/golden/ -- see tests/test_oracle_golden.py.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline/reference legs may load this library; the product never does.
//
// Every function cites the reference lines it restates.  Arithmetic follows the reference's
// operation order exactly (x86-64, no FMA contraction: built with -ffp-contract=off), because
// several results feed integer decisions (voxel index, sample level, Gauss set, z of a hit).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "tables_gen.h"

namespace {

struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, double k) { return {a.x * k, a.y * k, a.z * k}; }
inline V3 divs(V3 a, double k) { return {a.x / k, a.y / k, a.z / k}; }
// CGLA/ArithVec3Float.h:46-52
inline V3 cross(V3 x, V3 y) { return {x.y * y.z - x.z * y.y, x.z * y.x - x.x * y.z, x.x * y.y - x.y * y.x}; }
// CGLA/ArithVec.h:423-426: std::inner_product, left to right from T(0)
inline double dot(V3 a, V3 b) { return ((0.0 + a.x * b.x) + a.y * b.y) + a.z * b.z; }
inline double length(V3 a) { return std::sqrt(dot(a, a)); } // CGLA/ArithVecFloat.h:43-46
inline V3 normalize(V3 a) { return divs(a, length(a)); }    // CGLA/ArithVecFloat.h:74-77
inline V3 P(const double *pos, uint32_t n) { return {pos[3 * (size_t)n], pos[3 * (size_t)n + 1], pos[3 * (size_t)n + 2]}; }

// is_mesh/util.h:86-96
inline double signed_volume(V3 a, V3 b, V3 c, V3 d) { return dot(sub(a, d), cross(sub(b, d), sub(c, d))) / 6.; }
inline double volume(V3 a, V3 b, V3 c, V3 d) { return std::abs(signed_volume(a, b, c, d)); }
// is_mesh/util.h:80-84
inline double area(V3 a, V3 b, V3 c) { return 0.5 * length(cross(sub(b, a), sub(c, a))); }

struct Vol {
    int X, Y, Z, dtype; // 0 u8, 1 u16, 2 f32, 3 f64
    const void *data;
    // DEMO/image3d.cpp:480-491 (get_value) + the decode rule of DEMO/image3d.cpp:127
    inline double at(int x, int y, int z) const
    {
        if (x >= 0 && y >= 0 && z >= 0 && x < X && y < Y && z < Z) {
            size_t i = (size_t)z * X * Y + (size_t)y * X + x;
            switch (dtype) {
            case 0: return (double)((const uint8_t *)data)[i] / 255.0;
            case 1: return (double)((const uint16_t *)data)[i] / 65535.0;
            case 2: return (double)((const float *)data)[i];
            default: return ((const double *)data)[i];
            }
        }
        return 0.0; // BOUND_INTENSITY, DEMO/image3d.h:24
    }
    // DEMO/image3d.cpp:226-251 (get_value_f): floor, 8 fetches, lerp x then y then z
    inline double at_f(V3 pt) const
    {
        int ix = (int)std::floor(pt.x), iy = (int)std::floor(pt.y), iz = (int)std::floor(pt.z);
        double rx = pt.x - (double)ix, ry = pt.y - (double)iy, rz = pt.z - (double)iz;
        double c00 = at(ix, iy, iz) * (1 - rx) + at(ix + 1, iy, iz) * rx;
        double c01 = at(ix, iy, iz + 1) * (1 - rx) + at(ix + 1, iy, iz + 1) * rx;
        double c10 = at(ix, iy + 1, iz) * (1 - rx) + at(ix + 1, iy + 1, iz) * rx;
        double c11 = at(ix, iy + 1, iz + 1) * (1 - rx) + at(ix + 1, iy + 1, iz + 1) * rx;
        double c0 = c00 * (1 - ry) + c10 * ry;
        double c1 = c01 * (1 - ry) + c11 * ry;
        return c0 * (1 - rz) + c1 * rz;
    }
};

// sample level: DEMO/image3d.cpp:355-357 (upper_bound over {1,8,...,729}, minus one, clamped at 0)
inline int tet_level(double v)
{
    int idx = 0;
    while (idx < DSC_TET_LEVELS) {
        double sz = (double)(DSC_TET_LEVEL_OFFSET[idx + 1] - DSC_TET_LEVEL_OFFSET[idx]);
        if (v < sz) break;
        idx++;
    }
    int dis = idx - 1;
    return dis < 0 ? 0 : dis;
}

inline int first_not_in(const uint32_t *tn, const uint32_t *fn)
{   // (get_nodes(tid) - nids).front(), is_mesh/is_mesh.h:632-637
    for (int k = 0; k < 4; k++)
        if (tn[k] != fn[0] && tn[k] != fn[1] && tn[k] != fn[2]) return k;
    return 0;
}

// DSC::get_normal(fid, tid) = normal_direction(get_pos(get_sorted_nodes(fid, tid)))
// (src/DSC.h:3120-3124, is_mesh/is_mesh.h:632-637, 981-1000, util.h:366-376)
inline V3 face_normal_wrt(const double *pos, const uint32_t *fn, const uint32_t *tn)
{
    uint32_t s[3] = {fn[0], fn[1], fn[2]};
    uint32_t apex = tn[first_not_in(tn, fn)];
    V3 x = sub(P(pos, apex), P(pos, s[0]));
    V3 y = sub(P(pos, s[1]), P(pos, s[0]));
    V3 z = sub(P(pos, s[2]), P(pos, s[0]));
    if (dot(x, cross(y, z)) > 0.) std::swap(s[0], s[1]);
    V3 a = P(pos, s[0]), b = P(pos, s[1]), c = P(pos, s[2]);
    return normalize(cross(sub(b, a), sub(c, a)));
}

inline int gauss_set(double a)
{   // DEMO/gaussian_points.h:97-101; area >= 999999 would index past the 7 sets in the
    // reference (latent UB) -- clamped to the last set here and in the product.
    int r = 0;
    while (r < DSC_GAUSS_SETS && !(a < DSC_GAUSS_AREA_THRES[r])) r++;
    return r < DSC_GAUSS_SETS ? r : DSC_GAUSS_SETS - 1;
}

struct Hit { int z; bool b_in; };
inline bool hit_less(Hit a, Hit b) { return a.z < b.z; } // DEMO/segment_function.cpp:1595-1598

} // namespace

extern "C" {

int fo_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Per-tet sampling: DEMO/image3d.cpp:351-426.
//   total[t] = (sum I) * v / n3   (get_tetra_intensity's *total_inten)
//   vol[t]   = v ; mean[t] = total/v
//   sq[t*nc+k]  = (sum (I-c_k)^2) * v / n3   (get_energy)
//   ad[t*nc+k]  = sum |I-c_k|                (get_variation, not volume-normalised)
//   hash[t]  = FNV-1a over the truncated (x,y,z) of every sample, nsamp[t] = n3
// Tets with label 0 are processed too when tet_mask == NULL; otherwise tet_mask[t]==0 skips
// (outputs left untouched).  threads <= 1 runs the exact sequential order.
int fo_tet_integrate(const int dims[3], int dtype, const void *voxels, uint32_t n_tets, const uint32_t *tet_nodes, const double *pos,
                     const uint8_t *tet_mask, int nc, const double *c, double *total, double *vol, double *mean, double *sq, double *ad,
                     uint64_t *hash, int32_t *nsamp, int threads)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t t = 0; t < (int64_t)n_tets; t++) {
        if (tet_mask && !tet_mask[t]) continue;
        const uint32_t *tn = tet_nodes + 4 * t;
        V3 p0 = P(pos, tn[0]), p1 = P(pos, tn[1]), p2 = P(pos, tn[2]), p3 = P(pos, tn[3]);
        double v = volume(p0, p1, p2, p3);
        int lv = tet_level(v);
        int o = DSC_TET_LEVEL_OFFSET[lv], n3 = DSC_TET_LEVEL_OFFSET[lv + 1] - o;
        double s = 0;
        uint64_t h = 1469598103934665603ull;
        double sqa[16], ada[16];
        for (int k = 0; k < nc; k++) sqa[k] = ada[k] = 0;
        for (int i = 0; i < n3; i++) {
            const double *b = DSC_TET_BARY + 4 * (o + i);
            // get_coord: a[0]*b[0] + a[1]*b[1] + a[2]*b[2] + a[3]*b[3], DEMO/tet_dis_coord.hpp:27
            V3 pt = add(add(add(mul(p0, b[0]), mul(p1, b[1])), mul(p2, b[2])), mul(p3, b[3]));
            int x = (int)pt.x, y = (int)pt.y, z = (int)pt.z; // double -> const int truncation at the call, image3d.cpp:366
            double I = V.at(x, y, z);
            s += I;
            for (int k = 0; k < nc; k++) { sqa[k] += (I - c[k]) * (I - c[k]); ada[k] += std::abs(I - c[k]); }
            h ^= (uint64_t)(uint32_t)x; h *= 1099511628211ull;
            h ^= (uint64_t)(uint32_t)y; h *= 1099511628211ull;
            h ^= (uint64_t)(uint32_t)z; h *= 1099511628211ull;
        }
        double tot = s * v / (double)n3;
        if (total) total[t] = tot;
        if (vol) vol[t] = v;
        if (mean) mean[t] = tot / v;
        for (int k = 0; k < nc; k++) {
            if (sq) sq[t * (size_t)nc + k] = sqa[k] * v / (double)n3;
            if (ad) ad[t * (size_t)nc + k] = ada[k];
        }
        if (hash) hash[t] = h;
        if (nsamp) nsamp[t] = n3;
    }
    return 0;
}

// Per-phase reduction of per-tet sums in ascending tet order (the meaning of _total_intensities /
// _phase_volume per DEMO/segment_function.h:129-131); label 0 (BOUND_LABEL) is skipped as in
// DEMO/segment_function.cpp:165-168.
int fo_phase_reduce(uint32_t n_tets, const int32_t *tet_label, const double *tet_total, const double *tet_vol, int nb_phase,
                    double *total, double *volume, double *mean)
{
    for (int k = 0; k < nb_phase; k++) total[k] = volume[k] = 0;
    for (uint32_t t = 0; t < n_tets; t++) {
        int l = tet_label[t];
        if (l <= 0 || l > nb_phase) continue;
        total[l - 1] += tet_total[t];
        volume[l - 1] += tet_vol[t];
    }
    for (int k = 0; k < nb_phase; k++) mean[k] = total[k] / volume[k];
    return 0;
}

// External force: DEMO/segment_function.cpp:1425-1473.  forces is 3*n_nodes_buffer, zeroed here.
// Faces with a missing second tet (face_tets == 0xFFFFFFFF) are skipped (the reference would read
// tets[1] out of range; update_vertex_boundary prevents that case).
int fo_face_forces(const int dims[3], int dtype, const void *voxels, uint32_t n_nodes_buffer, const double *pos, const uint32_t *tet_nodes,
                   const int32_t *tet_label, uint32_t n_faces, const uint32_t *face_nodes, const uint32_t *face_tets, int nb_phase,
                   const double *mean, double *forces, int threads)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
    memset(forces, 0, sizeof(double) * 3 * (size_t)n_nodes_buffer);
    int T = threads > 1 ? threads : 1;
    std::vector<std::vector<double>> priv;
    if (T > 1) priv.assign(T, std::vector<double>(3 * (size_t)n_nodes_buffer, 0.0));
#pragma omp parallel for schedule(static) num_threads(T)
    for (int64_t f = 0; f < (int64_t)n_faces; f++) {
        double *F = forces;
#ifdef _OPENMP
        if (T > 1) F = priv[omp_get_thread_num()].data();
#endif
        uint32_t t0 = face_tets[2 * f], t1 = face_tets[2 * f + 1];
        if (t0 == 0xFFFFFFFFu || t1 == 0xFFFFFFFFu) continue;
        const uint32_t *fn = face_nodes + 3 * f;
        V3 pts[3] = {P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2])};
        int phase0 = tet_label[t0] - 1, phase1 = tet_label[t1] - 1;
        // phase_intensity, DEMO/segment_function.h:149-156
        double c0 = phase0 >= 0 ? mean[phase0] : -0.3, c1 = phase1 >= 0 ? mean[phase1] : -0.3;
        // get_normal(fid): sorted w.r.t. the co-boundary tet with the highest label, first wins
        // (is_mesh/is_mesh.h:639-663)
        uint32_t tid = tet_label[t1] > tet_label[t0] ? t1 : t0;
        V3 Norm = face_normal_wrt(pos, fn, tet_nodes + 4 * (size_t)tid);
        // barycenter(t) = (a+b+c+d)*0.25 in get_nodes(t) order, src/DSC.h:3254-3258, util.h:100-103
        const uint32_t *a = tet_nodes + 4 * (size_t)t0, *b = tet_nodes + 4 * (size_t)t1;
        V3 b0 = mul(add(add(add(P(pos, a[0]), P(pos, a[1])), P(pos, a[2])), P(pos, a[3])), 0.25);
        V3 b1 = mul(add(add(add(P(pos, b[0]), P(pos, b[1])), P(pos, b[2])), P(pos, b[3])), 0.25);
        V3 l01 = sub(b1, b0);
        Norm = mul(Norm, dot(Norm, l01));
        Norm = divs(Norm, length(Norm)); // Vec3d::normalize, CGLA/ArithVecFloat.h:49-52
        double ar = area(pts[0], pts[1], pts[2]);
        int gs = gauss_set(ar);
        for (int g = DSC_GAUSS_SET_OFFSET[gs]; g < DSC_GAUSS_SET_OFFSET[gs + 1]; g++) {
            const double *gp = DSC_GAUSS + 4 * g;
            double w = gp[0];
            V3 p = add(add(mul(pts[0], gp[1]), mul(pts[1], gp[2])), mul(pts[2], gp[3])); // get_coord_tri
            double I = V.at_f(p);
            V3 fv = mul(Norm, (2 * I - c0 - c1) * (c0 - c1) * w * ar);
            for (int ii = 0; ii < 3; ii++) {
                V3 d = mul(fv, gp[1 + ii]);
                double *o = F + 3 * (size_t)fn[ii];
                o[0] += d.x; o[1] += d.y; o[2] += d.z;
            }
        }
    }
    if (T > 1)
        for (int t = 0; t < T; t++)
            for (size_t i = 0; i < 3 * (size_t)n_nodes_buffer; i++) forces[i] += priv[t][i];
    return 0;
}

// Force loop of segment_function::thickenning_surface, DEMO/segment_function.cpp:1990-2051: compute_external_force's quadrature over the
// interface faces that are not segment_function::is_boundary(FaceKey) (all three nodes in m_vertex_bound, :2186-2196), and per face
// the mean and the mean absolute deviation of the magnitudes of its Gauss points (:2030-2043).  One thread (the reference's order).
int fo_thicken_forces(const int dims[3], int dtype, const void *voxels, uint32_t n_nodes_buffer, const double *pos, const uint32_t *tet_nodes,
                      const int32_t *tet_label, uint32_t n_faces, const uint32_t *face_nodes, const uint32_t *face_tets, const uint8_t *vb, int nb_phase,
                      const double *mean, double *forces, double *face_mean, double *face_dev, uint8_t *face_evaluated)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
    memset(forces, 0, sizeof(double) * 3 * (size_t)n_nodes_buffer);
    for (int64_t f = 0; f < (int64_t)n_faces; f++) {
        face_mean[f] = face_dev[f] = 0; face_evaluated[f] = 0;
        uint32_t t0 = face_tets[2 * f], t1 = face_tets[2 * f + 1];
        if (t0 == 0xFFFFFFFFu || t1 == 0xFFFFFFFFu) continue;
        const uint32_t *fn = face_nodes + 3 * f;
        if (vb[fn[0]] && vb[fn[1]] && vb[fn[2]]) continue;
        V3 pts[3] = {P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2])};
        int phase0 = tet_label[t0] - 1, phase1 = tet_label[t1] - 1;
        double c0 = phase0 >= 0 ? mean[phase0] : -0.3, c1 = phase1 >= 0 ? mean[phase1] : -0.3;
        uint32_t tid = tet_label[t1] > tet_label[t0] ? t1 : t0;
        V3 Norm = face_normal_wrt(pos, fn, tet_nodes + 4 * (size_t)tid);
        const uint32_t *a = tet_nodes + 4 * (size_t)t0, *b = tet_nodes + 4 * (size_t)t1;
        V3 b0 = mul(add(add(add(P(pos, a[0]), P(pos, a[1])), P(pos, a[2])), P(pos, a[3])), 0.25);
        V3 b1 = mul(add(add(add(P(pos, b[0]), P(pos, b[1])), P(pos, b[2])), P(pos, b[3])), 0.25);
        Norm = mul(Norm, dot(Norm, sub(b1, b0)));
        Norm = divs(Norm, length(Norm));
        double ar = area(pts[0], pts[1], pts[2]);
        int gs = gauss_set(ar);
        std::vector<double> mags;
        for (int g = DSC_GAUSS_SET_OFFSET[gs]; g < DSC_GAUSS_SET_OFFSET[gs + 1]; g++) {
            const double *gp = DSC_GAUSS + 4 * g;
            V3 p = add(add(mul(pts[0], gp[1]), mul(pts[1], gp[2])), mul(pts[2], gp[3]));
            double I = V.at_f(p);
            double magnitude = (2 * I - c0 - c1) * (c0 - c1) * gp[0] * ar;
            V3 fv = mul(Norm, magnitude);
            for (int ii = 0; ii < 3; ii++) {
                V3 d = mul(fv, gp[1 + ii]);
                double *o = forces + 3 * (size_t)fn[ii];
                o[0] += d.x; o[1] += d.y; o[2] += d.z;
            }
            mags.push_back(magnitude);
        }
        double mu = 0;
        for (double m : mags) mu += m;
        mu /= mags.size();
        double dev = 0;
        for (double m : mags) dev += std::abs(m - mu);
        dev /= mags.size();
        face_mean[f] = mu; face_dev[f] = dev; face_evaluated[f] = 1;
    }
    return 0;
}

// image3d::build_sum_table, DEMO/image3d.cpp:166-176: S[x,y,z] = S[x,y,z-1] + V[x,y,z].
// segment_function::compute_internal_force_2 (DEMO/segment_function.cpp:2197-2232).
//   pass 1, interface faces in ascending key order that are not all-vertex-bound (segment_function::is_boundary(FaceKey),
//           :2186-2196): corner i receives -h * |l2 - l1| with h = normalize(l0 - project_point_line(l0, l1, l2 - l1))
//           (is_mesh/util.h:395-399), l0/l1/l2 = corners i, i+1, i+2 in get_nodes(f) order;
//   pass 2, nodes that are interface and not crossing (node_flags bit 0 / bit 1, is_mesh/attributes.h:84-102) with
//           |f| > EPSILON (1e-8, is_mesh/util.h:43): f = n * dot(n, f), n = DSC::get_normal(node) (src/DSC.h:3138-3159):
//           the normalised sum of get_normal(face) over the node's interface faces, 0 when that sum is shorter than EPSILON.
// The reference adds the face normals in get_faces(node) order, which a flat snapshot does not carry.  order_off / order_idx
// (may be NULL) hand it in: the interface faces of get_faces(node) for node n are order_idx[order_off[n] .. order_off[n+1]) as indices
// into the face arrays -- then every force is bit-identical to the reference's.  Without them: ascending face order; the projected
// forces agree to rounding (1e-12), the unprojected sums bit for bit.
int fo_internal_forces(uint32_t n_nodes_buffer, const double *pos, const uint32_t *tet_nodes, const int32_t *tet_label, uint32_t n_faces,
                       const uint32_t *face_nodes, const uint32_t *face_tets, const uint8_t *vertex_bound, const uint8_t *node_flags,
                       const uint32_t *order_off, const uint32_t *order_idx, double *internal_forces)
{
    const double EPS = 1e-8;
    memset(internal_forces, 0, sizeof(double) * 3 * (size_t)n_nodes_buffer);
    std::vector<double> nsum(3 * (size_t)n_nodes_buffer, 0.0), fnorm(order_off ? 3 * (size_t)n_faces : 0, 0.0);
    for (uint32_t f = 0; f < n_faces; f++) {
        const uint32_t *fn = face_nodes + 3 * (size_t)f;
        uint32_t t0 = face_tets[2 * f], t1 = face_tets[2 * f + 1];
        // get_normal(fid): sorted w.r.t. the co-boundary tet with the highest label, first wins (is_mesh/is_mesh.h:639-663)
        uint32_t tid = t0;
        if (t0 == 0xFFFFFFFFu) tid = t1;
        else if (t1 != 0xFFFFFFFFu && tet_label[t1] > tet_label[t0]) tid = t1;
        if (tid != 0xFFFFFFFFu) {
            V3 N = face_normal_wrt(pos, fn, tet_nodes + 4 * (size_t)tid);
            for (int k = 0; k < 3; k++) { double *o = nsum.data() + 3 * (size_t)fn[k]; o[0] += N.x; o[1] += N.y; o[2] += N.z; }
            if (order_off) { fnorm[3 * (size_t)f] = N.x; fnorm[3 * (size_t)f + 1] = N.y; fnorm[3 * (size_t)f + 2] = N.z; }
        }
        if (vertex_bound[fn[0]] && vertex_bound[fn[1]] && vertex_bound[fn[2]]) continue;
        V3 p[3] = {P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2])};
        for (int i = 0; i < 3; i++) {
            V3 l0 = p[i], l1 = p[(i + 1) % 3], l2 = p[(i + 2) % 3];
            V3 v = sub(l2, l1), v1 = sub(l0, l1);
            V3 proj = add(l1, divs(mul(v, dot(v1, v)), dot(v, v))); // a + v * dot(v1,v) / dot(v,v)
            V3 h = normalize(sub(l0, proj));
            V3 t = mul(mul(h, -1.0), length(v)); // -h * (l2 - l1).length()
            double *o = internal_forces + 3 * (size_t)fn[i];
            o[0] += t.x; o[1] += t.y; o[2] += t.z;
        }
    }
    for (uint32_t n = 0; n < n_nodes_buffer; n++) {
        if (!(node_flags[n] & 1) || (node_flags[n] & 2)) continue;
        double *o = internal_forces + 3 * (size_t)n;
        V3 F{o[0], o[1], o[2]};
        if (!(length(F) > EPS)) continue;
        V3 S{nsum[3 * (size_t)n], nsum[3 * (size_t)n + 1], nsum[3 * (size_t)n + 2]};
        if (order_off) {
            S = V3{0, 0, 0};
            for (uint32_t k = order_off[n]; k < order_off[n + 1]; k++) {
                const double *q = fnorm.data() + 3 * (size_t)order_idx[k];
                S = add(S, V3{q[0], q[1], q[2]});
            }
        }
        V3 norm{0, 0, 0};
        if (!(length(S) < EPS)) norm = normalize(S);
        V3 r = mul(norm, dot(norm, F));
        o[0] = r.x; o[1] = r.y; o[2] = r.z;
    }
    return 0;
}

int fo_build_sum_table(const int dims[3], int dtype, const void *voxels, double *table)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
    size_t XY = (size_t)dims[0] * dims[1];
    for (int y = 0; y < dims[1]; y++)
        for (int x = 0; x < dims[0]; x++) {
            size_t i = (size_t)y * dims[0] + x;
            table[i] = V.at(x, y, 0);
            for (int z = 1; z < dims[2]; z++) table[i + z * XY] = table[i + (z - 1) * XY] + V.at(x, y, z);
        }
    return 0;
}

// m_vertex_bound of segment_function::update_vertex_boundary (DEMO/segment_function.cpp:2158-2171),
// evaluated on a snapshot taken after the relabelling part (:2148-2155) has run on the host.
int fo_vertex_bound(uint32_t n_nodes_buffer, const int32_t *tet_label, uint32_t n_faces, const uint32_t *face_nodes,
                    const uint32_t *face_tets, uint8_t *vertex_bound)
{
    memset(vertex_bound, 0, n_nodes_buffer);
    for (uint32_t f = 0; f < n_faces; f++) {
        uint32_t t0 = face_tets[2 * f], t1 = face_tets[2 * f + 1];
        if (t0 == 0xFFFFFFFFu || t1 == 0xFFFFFFFFu) continue;
        if (tet_label[t0] == 0 || tet_label[t1] == 0)
            for (int k = 0; k < 3; k++) vertex_bound[face_nodes[3 * f + k]] = 1;
    }
    return 0;
}

// z-ray pass shared by update_average_intensity (DEMO/segment_function.cpp:1600-1764) and
// compute_energy (:2331-2492).
//   mode 0: total[k] += S[z2]-S[z1] via sum_line_z (image3d.cpp:179-205), count[k] += z2-z1, mean = total/count
//   mode 1: energy_data += (I(x,y,k)-c)^2 for k in (z1,z2) exclusive (:2470-2473);
//           energy_area  = sum alpha*area over interface faces whose three nodes are not all vertex-bound (:2482-2489)
// hits_hist (optional, 64 ints) receives the histogram of hits per kept ray, n_rays the number of kept rays.
int fo_zray(const int dims[3], int dtype, const void *voxels, const double *sum_table, const double *pos, const uint32_t *tet_nodes,
            const int32_t *tet_label, uint32_t n_faces, const uint32_t *face_nodes, const uint32_t *face_tets, const uint8_t *face_is_boundary,
            int nb_phase, int mode, const double *mean_in, double alpha, const uint8_t *vertex_bound, double *total, double *count,
            double *mean_out, double *energy_data, double *energy_area, int64_t *n_rays, int32_t *hits_hist)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
    const int X = dims[0], Y = dims[1], Z = dims[2];
    size_t XY = (size_t)X * Y;
    std::vector<std::vector<Hit>> rays((size_t)nb_phase * XY);
    for (uint32_t f = 0; f < n_faces; f++) {
        if (face_is_boundary[f]) continue; // is_interface() and !is_boundary(), :1628
        uint32_t t0 = face_tets[2 * f], t1 = face_tets[2 * f + 1];
        if (t0 == 0xFFFFFFFFu || t1 == 0xFFFFFFFFu) continue;
        int phase0 = tet_label[t0] - 1, phase1 = tet_label[t1] - 1;
        const uint32_t *fn = face_nodes + 3 * f;
        V3 p3[3] = {P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2])};
        V3 p2[3] = {{p3[0].x, p3[0].y, 0}, {p3[1].x, p3[1].y, 0}, {p3[2].x, p3[2].y, 0}};
        V3 n = face_normal_wrt(pos, fn, tet_nodes + 4 * (size_t)t0);
        bool in_1 = dot(n, V3{0, 0, 1}) > 0;
        double ldx = HUGE_VAL, ldy = HUGE_VAL, rux = -HUGE_VAL, ruy = -HUGE_VAL; // bounding_box, :1578-1593
        for (int k = 0; k < 3; k++) {
            ldx = std::min(p2[k].x, ldx); ldy = std::min(p2[k].y, ldy);
            rux = std::max(p2[k].x, rux); ruy = std::max(p2[k].y, ruy);
        }
        for (int x = (int)std::floor(ldx); x < std::round(rux); x++)
            for (int y = (int)std::floor(ldy); y < std::round(ruy); y++) {
                if (y < 0 || x < 0 || y >= Y || x >= X) continue;
                // 2-D barycentric_coords with error flag, is_mesh/util.h:283-315
                V3 p{x + 0.5, y + 0.5, 0};
                V3 v0 = sub(p2[1], p2[0]), v1 = sub(p2[2], p2[0]), v2 = sub(p, p2[0]);
                double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
                double denom = d00 * d11 - d01 * d01;
                if (std::abs(denom) < 1e-8) continue;
                double bc1 = (d11 * d20 - d01 * d21) / denom;
                double bc2 = (d00 * d21 - d01 * d20) / denom;
                double bc0 = 1. - bc1 - bc2;
                if (bc0 > -1e-8 && bc1 > -1e-8 && bc2 > -1e-8) {
                    double pz = (p3[0].z * bc0 + p3[1].z * bc1) + p3[2].z * bc2;
                    int zz = (int)std::nearbyint(pz);
                    if (phase0 != -1) rays[(size_t)phase0 * XY + (size_t)y * X + x].push_back({zz, !in_1});
                    if (phase1 != -1) rays[(size_t)phase1 * XY + (size_t)y * X + x].push_back({zz, in_1});
                }
            }
    }
    if (hits_hist) memset(hits_hist, 0, sizeof(int32_t) * 64);
    int64_t kept = 0;
    double diff = 0;
    for (int i = 0; i < nb_phase; i++) {
        double tot = 0, cnt = 0;
        double cc = mean_in ? mean_in[i] : 0;
        for (size_t r = 0; r < XY; r++) {
            std::vector<Hit> ps = rays[(size_t)i * XY + r];
            if (ps.size() <= 1) continue;
            int rx = (int)(r % X), ry = (int)(r / X);
            std::sort(ps.begin(), ps.end(), hit_less);
            // adjacent-duplicate removal with the reference's skip-after-erase quirk, :1706-1719
            for (auto p = ps.begin() + 1; p != ps.end(); p++) {
                auto pre = p - 1;
                if (p->z == pre->z && p->b_in == pre->b_in) {
                    p = ps.erase(p);
                    if (p == ps.end()) break;
                }
            }
            if (ps.size() % 2 != 0) continue; // odd rays silently dropped, :1721,1741
            kept++;
            if (hits_hist) hits_hist[std::min<size_t>(ps.size(), 63)]++;
            for (size_t j = 0; j < ps.size() / 2; j++) {
                int z1 = ps[2 * j].z, z2 = ps[2 * j + 1].z;
                if (z1 < 0) z1 = 0;
                if (z2 < z1) z2 = z1;
                cnt += z2 - z1;
                if (mode == 0) {
                    int a = z1, b = z2; // sum_line_z clamps, image3d.cpp:193-201
                    if (b >= Z) { b = Z - 1; if (a >= Z) a = Z - 1; }
                    tot += sum_table[(size_t)b * XY + (size_t)ry * X + rx] - sum_table[(size_t)a * XY + (size_t)ry * X + rx];
                } else {
                    for (int k = z1 + 1; k < z2; k++) { double d = V.at(rx, ry, k) - cc; diff += d * d; }
                }
            }
        }
        if (total) total[i] = tot;
        if (count) count[i] = cnt;
        if (mean_out) mean_out[i] = tot / cnt;
    }
    if (n_rays) *n_rays = kept;
    if (mode == 1) {
        double areaa = 0;
        for (uint32_t f = 0; f < n_faces; f++) {
            const uint32_t *fn = face_nodes + 3 * f;
            bool allb = vertex_bound[fn[0]] && vertex_bound[fn[1]] && vertex_bound[fn[2]]; // segment_function::is_boundary(FaceKey), :2186-2196
            if (!allb) areaa += alpha * area(P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2]));
        }
        if (energy_data) *energy_data = diff;
        if (energy_area) *energy_area = areaa;
    }
    return 0;
}

// 4-point barycentric "voxel-in-tet" predicate: is_point_inside, DEMO/segment_function.cpp:2576-2580,
// on Util::barycentric_coords (is_mesh/util.h:349-364).
int fo_points_in_tets(uint32_t n, const uint32_t *tet_idx, const double *pts, const uint32_t *tet_nodes, const double *pos, uint8_t *inside)
{
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t *tn = tet_nodes + 4 * (size_t)tet_idx[i];
        V3 a = P(pos, tn[0]), b = P(pos, tn[1]), c = P(pos, tn[2]), d = P(pos, tn[3]);
        V3 p{pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
        double co[4] = {signed_volume(p, b, c, d), signed_volume(a, p, c, d), signed_volume(a, b, p, d), signed_volume(a, b, c, p)};
        double s = co[0] + co[1] + co[2] + co[3];
        for (int k = 0; k < 4; k++) co[k] /= s;
        inside[i] = (co[0] >= 0 && co[1] >= 0 && co[2] >= 0 && co[3] >= 0) ? 1 : 0;
    }
    return 0;
}

// Relabel energy of every labelled tet for every candidate label: segment_function::get_energy_tetrahedron
// (DEMO/segment_function.cpp:618-643) as called by relabel_tetrahedra (:845-890) with the labels left untouched:
//   energy[t][k] = get_variation(t, mean[k]) + sum over the 4 faces in get_faces(t) order of alpha * area(face) when the
//   "other" label differs from k+1, where other = (label(tets[0]) == old ? label(tets[1]) : label(tets[0])) (:633-636).
// tet_face_nodes [n_tets][4][3] in get_nodes(f) order, tet_face_tets [n_tets][4][2] in get_tets(f) order (tet indices).
// Faces with a single co-boundary tet are skipped (the reference assumes there are none, comment at :632).
int fo_relabel_energies(const int dims[3], int dtype, const void *voxels, uint32_t n_tets, const uint32_t *tet_nodes, const int32_t *tet_label,
                        const double *pos, const uint32_t *tet_face_nodes, const uint32_t *tet_face_tets, int nb_phase, const double *mean, double alpha,
                        double *energy)
{
    Vol V{dims[0], dims[1], dims[2], dtype, voxels};
    for (uint32_t t = 0; t < n_tets; t++) {
        for (int k = 0; k < nb_phase; k++) energy[(size_t)t * nb_phase + k] = -1.0;
        const int old_label = tet_label[t];
        if (old_label == 0) continue;
        const uint32_t *tn = tet_nodes + 4 * (size_t)t;
        V3 p0 = P(pos, tn[0]), p1 = P(pos, tn[1]), p2 = P(pos, tn[2]), p3 = P(pos, tn[3]);
        double v = volume(p0, p1, p2, p3);
        int lv = tet_level(v);
        int o = DSC_TET_LEVEL_OFFSET[lv], n3 = DSC_TET_LEVEL_OFFSET[lv + 1] - o;
        for (int k = 0; k < nb_phase; k++) {
            double e = 0;
            for (int i = 0; i < n3; i++) { // get_variation, DEMO/image3d.cpp:406-426
                const double *b = DSC_TET_BARY + 4 * (o + i);
                V3 pt = add(add(add(mul(p0, b[0]), mul(p1, b[1])), mul(p2, b[2])), mul(p3, b[3]));
                e += std::abs(V.at((int)pt.x, (int)pt.y, (int)pt.z) - mean[k]);
            }
            for (int f = 0; f < 4; f++) {
                const uint32_t ta = tet_face_tets[((size_t)t * 4 + f) * 2], tb = tet_face_tets[((size_t)t * 4 + f) * 2 + 1];
                if (ta == 0xFFFFFFFFu || tb == 0xFFFFFFFFu) continue;
                int label0 = tet_label[ta], label1 = tet_label[tb];
                label0 = label0 == old_label ? label1 : label0;
                if (label0 != k + 1) {
                    const uint32_t *fn = tet_face_nodes + ((size_t)t * 4 + f) * 3;
                    e += area(P(pos, fn[0]), P(pos, fn[1]), P(pos, fn[2])) * alpha;
                }
            }
            energy[(size_t)t * nb_phase + k] = e;
        }
    }
    return 0;
}

// reference sort used by the z-ray pass, exposed so that the product's device-side restatement of
// libstdc++'s introsort can be fuzzed against the real std::sort (tests/test_introsort.py).
int fo_std_sort_hits(int n, int32_t *z, uint8_t *b_in)
{
    std::vector<Hit> h(n);
    for (int i = 0; i < n; i++) h[i] = {z[i], b_in[i] != 0};
    std::sort(h.begin(), h.end(), hit_less);
    for (int i = 0; i < n; i++) { z[i] = h[i].z; b_in[i] = h[i].b_in; }
    return 0;
}

} // extern "C"
