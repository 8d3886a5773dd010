// dsc_kernels.cuh -- sm_100a kernels of the DSC image-energy path.
//
// HBM layout (all arrays SoA, see DESIGN.md "Data layout"):
//   volume     VoxT[X*Y*Z]            x fastest (DEMO/image3d.h:66-68); u8/u16/f32/f64
//   sum_table  double[X*Y*Z]          z prefix sums (DEMO/image3d.cpp:166-176)
//   pos4       double4[n_nodes]       xyz + pad: one 32 B sector per node gather
//   tet_nodes  uint4[n_tets]          get_nodes(t) order, one 16 B load per tet
//   tet_label  int[n_tets]
//   face_nodes uint[3*n_faces], face_tets uint[2*n_faces], face_is_boundary u8[n_faces]
//   node CSR   node_off uint[n_nodes+1], node_inc uint[3*n_faces] = (face<<2)|corner, faces ascending
//   tet_bary   double[4*2025]         barycentric sample table (generated dsc_tables.h)
// None of this is GEMM-shaped; tensor cores are not used.  The kernels are gather/scan work
// bounded by HBM/L2 bandwidth and, for the tet pass, the FP64 pipe (21 non-fused DMUL/DADD per
// sample position, which must not be contracted -- see dsc_device.cuh).
#pragma once
#include <climits>
#include <type_traits>
#include "dsc_device.cuh"

#ifndef W2_MIN_BLOCKS
#define W2_MIN_BLOCKS 3
#endif

namespace dsc {

constexpr int kMaxPhase = 8;
constexpr int kMaxRayHits = 96;
constexpr unsigned kNoTet = 0xFFFFFFFFu;

__constant__ double c_gauss[4 * 42];
__constant__ int c_gauss_off[8];

struct VolumeView {
    const void *data;
    const double *sum_table;
    int X, Y, Z;
};

struct MeshView {
    const double4 *pos4;
    const float4 *pos4f; // float copy of pos4 (FP32 filter path), written by pos_to_float_kernel
    const uint4 *tet_nodes;
    const int *tet_label;
    const unsigned *face_nodes;
    const unsigned *face_tets;
    const unsigned char *face_is_boundary;
    const unsigned *node_off;
    const unsigned *node_inc;
    unsigned n_nodes, n_tets, n_faces;
};

template <typename T> struct VoxTraits;
template <> struct VoxTraits<unsigned char> { static constexpr bool integral = true; DSC_D static double denom() { return 255.0; } };
template <> struct VoxTraits<unsigned short> { static constexpr bool integral = true; DSC_D static double denom() { return 65535.0; } };
template <> struct VoxTraits<float> { static constexpr bool integral = false; DSC_D static double denom() { return 1.0; } };
template <> struct VoxTraits<double> { static constexpr bool integral = false; DSC_D static double denom() { return 1.0; } };

// exact voxel decode: byte/255.0 as DEMO/image3d.cpp:127 (IEEE division), u16/65535.0, float widened
template <typename T> DSC_D double decode(T v)
{
    if (VoxTraits<T>::integral) return (double)v / VoxTraits<T>::denom();
    return (double)v;
}

// image3d::get_value (DEMO/image3d.cpp:480-491): 0 outside the volume, never clamped
template <typename T> DSC_D double fetch(const T *__restrict__ vol, int X, int Y, int Z, int x, int y, int z)
{
    if ((unsigned)x < (unsigned)X && (unsigned)y < (unsigned)Y && (unsigned)z < (unsigned)Z)
        return decode<T>(__ldg(vol + ((size_t)z * Y + y) * X + x));
    return 0.0;
}

// branch-free form of fetch: the load is always issued (from a clamped address) and the result selected, so that
// independent fetches (the 8 corners of a trilinear cell) are in flight together instead of one per taken branch
template <typename T> DSC_D double fetch_nb(const T *__restrict__ vol, int X, int Y, int Z, int x, int y, int z)
{
    const bool inb = (unsigned)x < (unsigned)X && (unsigned)y < (unsigned)Y && (unsigned)z < (unsigned)Z;
    const size_t idx = inb ? ((size_t)z * Y + y) * X + x : 0;
    const T raw = __ldg(vol + idx);
    return inb ? decode<T>(raw) : 0.0;
}

// image3d::get_value_f (DEMO/image3d.cpp:226-251): floor, 8 fetches, lerp along x, then y, then z
template <typename T> DSC_D double fetch_trilinear(const T *__restrict__ vol, int X, int Y, int Z, V3 pt)
{
    double fx = floor(pt.x), fy = floor(pt.y), fz = floor(pt.z);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    double rx = pt.x - (double)ix, ry = pt.y - (double)iy, rz = pt.z - (double)iz;
    double v000 = fetch_nb(vol, X, Y, Z, ix, iy, iz), v100 = fetch_nb(vol, X, Y, Z, ix + 1, iy, iz);
    double v001 = fetch_nb(vol, X, Y, Z, ix, iy, iz + 1), v101 = fetch_nb(vol, X, Y, Z, ix + 1, iy, iz + 1);
    double v010 = fetch_nb(vol, X, Y, Z, ix, iy + 1, iz), v110 = fetch_nb(vol, X, Y, Z, ix + 1, iy + 1, iz);
    double v011 = fetch_nb(vol, X, Y, Z, ix, iy + 1, iz + 1), v111 = fetch_nb(vol, X, Y, Z, ix + 1, iy + 1, iz + 1);
    double c00 = v000 * (1 - rx) + v100 * rx;
    double c01 = v001 * (1 - rx) + v101 * rx;
    double c10 = v010 * (1 - rx) + v110 * rx;
    double c11 = v011 * (1 - rx) + v111 * rx;
    double c0 = c00 * (1 - ry) + c10 * ry;
    double c1 = c01 * (1 - ry) + c11 * ry;
    return c0 * (1 - rz) + c1 * rz;
}

// raw element fetch through a 3-D texture (cudaArray, block-linear layout, point sampling, border mode: reads
// outside the volume return 0 = BOUND_INTENSITY, exactly like image3d::get_value)
template <typename T> DSC_D T tex_fetch(cudaTextureObject_t tex, int x, int y, int z)
{
    return tex3D<T>(tex, (float)x + 0.5f, (float)y + 0.5f, (float)z + 0.5f);
}

// ---- bricked, padded copy of the volume (built by brick_volume_kernel, dsc_tet_g32.cuh) ----
// 4x2x4-voxel bricks (32 voxels; 128 B for float), padded by kBrickPad voxels on every side.  The padding holds what
// image3d::get_value returns outside the volume (0, DEMO/image3d.cpp:480-491) and, in the layer at -1, a replica of layer 0:
// the reference truncates toward zero (image3d.cpp:366), so a coordinate in (-1,0) reads voxel 0.  Element 0 is padding (0).
constexpr int kBrickPad = 8; // a multiple of every brick extent

// element index of voxel (x, y, z) (volume coordinates, -kBrickPad <= x < X + pad ...) in the bricked, padded copy
DSC_HD unsigned brick_index(int x, int y, int z, unsigned Sy, unsigned Sz)
{
    const unsigned xp = (unsigned)(x + kBrickPad), yp = (unsigned)(y + kBrickPad), zp = (unsigned)(z + kBrickPad);
    return (zp >> 2) * Sz + (yp >> 1) * Sy + ((xp >> 2) << 5) + ((zp & 3u) << 3) + ((yp & 1u) << 2) + (xp & 3u);
}

// Which tets a pass integrates: phases always, the boundary layer (label 0 = BOUND_LABEL, DEMO/image3d.h:23) when asked; a negative
// label (DSCGPU_LABEL_FREE) marks a slot of a key-indexed snapshot that holds no tet: never.
DSC_D bool tet_active(int label, int include_bound) { return label > 0 || (label == 0 && include_bound); }

DSC_D V3 load_pos(const double4 *__restrict__ pos4, unsigned n)
{
    const double2 *p = reinterpret_cast<const double2 *>(pos4 + n);
    double2 a = __ldg(p), b = __ldg(p + 1);
    return {a.x, a.y, b.x};
}

// Exact float -> double widening with integer ops.  F2F.F64.F32 runs on the XU pipe, which the FP64<->int/float
// conversions of this kernel saturate (ncu: xu_realtime 84 %); the ALU pipe is idle.  Normal numbers only need the
// exponent re-biased (+896) and the mantissa shifted; zeros, denormals, infinities and NaNs take the F2F path.
DSC_D double widen_exact(float f)
{
    const unsigned b = __float_as_uint(f);
    const unsigned e = b & 0x7f800000u;
    if (e == 0u || e == 0x7f800000u) return (double)f;
    const unsigned hi = (((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u);
    const unsigned lo = b << 29;
    return __hiloint2double((int)hi, (int)lo);
}

// A non-negative finite float with bits b is the double with words (b >> 3, b << 29) times 2^896 -- zeros and denormals
// included -- so sums of such values can be kept scaled by 2^-896: two shifts per voxel, no conversion instruction, no
// special cases, and one exact multiplication by 2^896 when the sum is used.  Volumes that hold negative values (incl. -0),
// infinities or NaNs (volume_flags_kernel) use widen_exact.
DSC_D double widen_scaled(float f)
{
    const unsigned b = __float_as_uint(f);
    return __hiloint2double((int)(b >> 3), (int)(b << 29));
}

// flags[0] bit 0: the float volume holds something other than non-negative finite numbers; flags[1]: the largest bit pattern
// (for non-negative finite floats: the bit pattern of the largest voxel)
__global__ void volume_flags_kernel(const float *__restrict__ vol, size_t n, unsigned *flags)
{
    unsigned f = 0, mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(__ldg(vol + i));
        if (b >= 0x7f800000u) f = 1;
        else mx = max(mx, b);
    }
    if (__any_sync(0xffffffffu, f) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(flags + 1, mx);
}

__global__ void pos_to_float_kernel(const double4 *__restrict__ pos4, float4 *__restrict__ pos4f, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 a = reinterpret_cast<const double2 *>(pos4 + i)[0], b = reinterpret_cast<const double2 *>(pos4 + i)[1];
    pos4f[i] = make_float4((float)a.x, (float)a.y, (float)b.x, 0.f);
}

DSC_D double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// =================================================================================================
// Per-tet integration: image3d::get_tetra_intensity / get_energy / get_variation
// (DEMO/image3d.cpp:351-426) over all tets, fused with the per-label reduction.
//
// G lanes cooperate on one tet (G = 8 or 32): all of them read the tet's 4 node ids and 4 positions
// (broadcast loads), compute volume -> level n -> n^3 sample points; lane j takes samples j, j+G,...
// from the flat barycentric table (coalesced 32 B rows), computes the position with the reference's
// operation order, truncates to the voxel index and gathers the voxel.  A G-wide xor-shuffle tree
// reduces the lane sums.  The per-label sums live in registers of lane (label-1) of each group, are
// reduced per CTA and written as one partial row per CTA; tet_finalize_kernel adds the rows in CTA
// order, so the result is bit-reproducible run to run (no atomics on doubles anywhere).
// =================================================================================================
struct alignas(64) TmapBlob { unsigned long long v[16]; }; // a CUtensorMap by value (kernel parameter space, __grid_constant__)

struct TetArgs {
    VolumeView vol;
    MeshView mesh;
    const double *tet_bary;
    unsigned tet_begin, tet_end;
    int include_bound;
    int nb_phase;
    int nc;
    double c[kMaxPhase];
    double *tet_total, *tet_vol, *tet_mean, *tet_energy, *tet_variation;
    const float4 *tet_bary_f;        // the same table rounded to float (FP32 filter path)
    const float4 *tet_lat;           // lattice form of the table: (a1, a2, a3, eps) per row, see tet_integrate_w2_kernel
    int want_sq;                     // fill the per-label second moment (costs one DFMA per sample)
    double *partials;                // [gridDim.x][3*kMaxPhase]: total, sumsq, volume
    unsigned long long *sample_count; // algorithmic samples processed (for the roofline figure)
    const void *brick;               // bricked, padded copy of the volume (w2 BRICK form), strides in elements
    unsigned brick_sy, brick_sz;
    int vol_simple;                  // float volume: only non-negative finite values (scaled widening allowed)
    int pack;                        // w2: packed FP32 position arithmetic
    int lean;                        // w2: full/tail batch split, per-tet weight from phase A
    // staged kernel (dsc_tet_staged.cuh)
    const TmapBlob *stg_maps;        // [16][16][16] tensor maps of the volume in device memory: box = (u+1) x 16 bytes by (r+1) rows by (s+1) slices
    int stg_slot_bytes;              // voxel box of one shared-memory slot (multiple of 1024)
    int stg_slots;                   // slots of the ring
    unsigned *stg_counter;           // tile fetch counter, zero at launch
    int stg_chunk;                   // tiles a producer warp claims at a time
    unsigned *stg_exact_count;       // tets left to the exact path (stg_counter + 1), zero at the first launch of a pass
    const unsigned *stg_exact_done;  // how many of them earlier tet_exact_list_kernel launches of this pass have worked off (stg_counter + 2)
    unsigned *stg_exact_list;        // ... and their indices, capacity tet_end - tet_begin
    unsigned long long *stg_acc;     // [3*kMaxPhase] fixed-point sums (total, sumsq, volume per label) the CTAs of both staged kernels add theirs to
    double stg_scale[3];             // fixed-point scales (powers of two) of total, sumsq, volume
};

template <typename VoxT, int G, bool WITH_C>
__global__ void __launch_bounds__(256) tet_integrate_kernel(const TetArgs a)
{
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    const int X = a.vol.X, Y = a.vol.Y, Z = a.vol.Z;
    const int lane_g = threadIdx.x & (G - 1);
    const unsigned groups_per_block = blockDim.x / G;
    const unsigned total_groups = gridDim.x * groups_per_block;
    const unsigned gid = blockIdx.x * groups_per_block + threadIdx.x / G;
    const unsigned n_work = a.tet_end - a.tet_begin;
    const unsigned iters = (n_work + total_groups - 1) / total_groups;

    double acc_tot = 0, acc_sq = 0, acc_vol = 0; // per-label sums of label == lane_g + 1
    unsigned long long nsamp = 0;

    for (unsigned it = 0; it < iters; it++) {
        const unsigned w = it * total_groups + gid;
        const bool valid = w < n_work;
        const unsigned t = a.tet_begin + (valid ? w : 0);
        int label = 0;
        bool active = false;
        double s = 0, sq = 0, v = 0;
        double sqc[kMaxPhase], adc[kMaxPhase];
        if (WITH_C) {
#pragma unroll
            for (int k = 0; k < kMaxPhase; k++) sqc[k] = adc[k] = 0;
        }
        int n3 = 0;
        if (valid) {
            label = __ldg(a.mesh.tet_label + t);
            active = tet_active(label, a.include_bound);
        }
        if (active) {
            const uint4 nd = __ldg(a.mesh.tet_nodes + t);
            const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
            const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
            v = tet_volume(p0, p1, p2, p3);
            const int L = tet_level(v);
            n3 = (L + 1) * (L + 1) * (L + 1);
            const int off = (L * (L + 1) / 2) * (L * (L + 1) / 2); // sum of k^3, k <= L
            const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(a.tet_bary) + 2 * (size_t)off;
            unsigned long long isum = 0, isq = 0;
            for (int i = lane_g; i < n3; i += G) {
                const double2 b01 = __ldg(tb + 2 * i), b23 = __ldg(tb + 2 * i + 1);
                // get_coord: ((p0*b0 + p1*b1) + p2*b2) + p3*b3 per component (DEMO/tet_dis_coord.hpp:27)
                const double px = ((p0.x * b01.x + p1.x * b01.y) + p2.x * b23.x) + p3.x * b23.y;
                const double py = ((p0.y * b01.x + p1.y * b01.y) + p2.y * b23.x) + p3.y * b23.y;
                const double pz = ((p0.z * b01.x + p1.z * b01.y) + p2.z * b23.x) + p3.z * b23.y;
                const int ix = (int)px, iy = (int)py, iz = (int)pz; // truncation toward 0, image3d.cpp:366
                if (VoxTraits<VoxT>::integral && !WITH_C) {
                    // integer voxels: exact integer sums, one division per tet instead of one per sample
                    if ((unsigned)ix < (unsigned)X && (unsigned)iy < (unsigned)Y && (unsigned)iz < (unsigned)Z) {
                        const unsigned long long q = (unsigned long long)__ldg(vol + ((size_t)iz * Y + iy) * X + ix);
                        isum += q;
                        isq += q * q;
                    }
                } else {
                    const double I = fetch(vol, X, Y, Z, ix, iy, iz);
                    s += I;
                    sq += I * I;
                    if (WITH_C) {
#pragma unroll
                        for (int k = 0; k < kMaxPhase; k++)
                            if (k < a.nc) {
                                const double d = I - a.c[k];
                                sqc[k] += d * d;
                                adc[k] += fabs(d);
                            }
                    }
                }
            }
            if (VoxTraits<VoxT>::integral && !WITH_C) {
                const double dn = VoxTraits<VoxT>::denom();
                s = (double)isum / dn;
                sq = (double)isq / (dn * dn);
            }
        }
        // group reduction (all 32 lanes take part; idle groups carry zeros)
#pragma unroll
        for (int m = G / 2; m >= 1; m >>= 1) {
            s += shfl_xor_d(s, m);
            sq += shfl_xor_d(sq, m);
            if (WITH_C) {
#pragma unroll
                for (int k = 0; k < kMaxPhase; k++) {
                    sqc[k] += shfl_xor_d(sqc[k], m);
                    adc[k] += shfl_xor_d(adc[k], m);
                }
            }
        }
        if (active) {
            // total = total * v / a.size()  (image3d.cpp:371)
            const double tot = s * v / (double)n3;
            const double tsq = sq * v / (double)n3;
            if (lane_g == 0) {
                nsamp += (unsigned long long)n3;
                if (a.tet_total) a.tet_total[t] = tot;
                if (a.tet_vol) a.tet_vol[t] = v;
                if (a.tet_mean) a.tet_mean[t] = tot / v;
                if (WITH_C) {
#pragma unroll
                    for (int k = 0; k < kMaxPhase; k++)
                        if (k < a.nc) {
                            if (a.tet_energy) a.tet_energy[(size_t)t * a.nc + k] = sqc[k] * v / (double)n3;
                            if (a.tet_variation) a.tet_variation[(size_t)t * a.nc + k] = adc[k];
                        }
                }
            }
            if (label >= 1 && label <= a.nb_phase && lane_g == label - 1) {
                acc_tot += tot;
                acc_sq += tsq;
                acc_vol += v;
            }
        } else if (valid && lane_g == 0) {
            if (a.tet_total) a.tet_total[t] = -1.0;
            if (a.tet_vol) a.tet_vol[t] = -1.0;
            if (a.tet_mean) a.tet_mean[t] = -1.0;
            if (WITH_C)
                for (int k = 0; k < a.nc; k++) {
                    if (a.tet_energy) a.tet_energy[(size_t)t * a.nc + k] = -1.0;
                    if (a.tet_variation) a.tet_variation[(size_t)t * a.nc + k] = -1.0;
                }
        }
    }

    // ---- CTA reduction of the per-label sums: lanes with equal (lane & (G-1)) hold the same label
#pragma unroll
    for (int m = 16; m >= G; m >>= 1) {
        acc_tot += shfl_xor_d(acc_tot, m);
        acc_sq += shfl_xor_d(acc_sq, m);
        acc_vol += shfl_xor_d(acc_vol, m);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, m);
    __shared__ double sm[8][3 * kMaxPhase];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < kMaxPhase && lane < G) {
        sm[warp][lane] = acc_tot;
        sm[warp][kMaxPhase + lane] = acc_sq;
        sm[warp][2 * kMaxPhase + lane] = acc_vol;
    }
    __syncthreads();
    if (threadIdx.x < 3 * kMaxPhase) {
        double r = 0;
        const int nw = blockDim.x >> 5;
        for (int wdx = 0; wdx < nw; wdx++) r += sm[wdx][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 3 * kMaxPhase + threadIdx.x] = r;
    }
    if (lane == 0 && nsamp) atomicAdd(a.sample_count, nsamp);
}

// levels whose table rows the w2 kernel stages in shared memory
constexpr int kSmemLevels = 5;                                                        // n = 1..5
constexpr int kSmemRows = (kSmemLevels * (kSmemLevels + 1) / 2) * (kSmemLevels * (kSmemLevels + 1) / 2); // 225 rows

// ---- packed FP32 pairs (Blackwell FFMA2 / FADD2, PTX fma.rn.f32x2 / add.rn.f32x2): two IEEE operations per issue slot ----
typedef unsigned long long f32x2;
DSC_D f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
DSC_D void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DSC_D f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
DSC_D f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// first 32-row slot of level L (n = L + 1) when every level is padded to a multiple of 32 rows
DSC_HD int g32_slot_off(int L)
{
    // {0, 1, 2, 3, 5, 9, 16, 27, 43, 66} packed in bytes (a local array would live on the stack)
    const unsigned w = L < 4 ? 0x03020100u : (L < 8 ? 0x1B100905u : 0x0000422Bu);
    return (int)((w >> (8 * (L & 3))) & 0xffu);
}

// -------------------------------------------------------------------------------------------------
// Warp-level two-phase form with origin-relative lattice coordinates (variant 128).
//
// A warp takes a tile of 32 consecutive tets.
//   phase A  lane <-> tet: label, ids, FP64 positions -> exact volume and sample level (as the reference), then the
//            record of the FP32 filter: integer origin O = floor(min corner), Q_k = P_k - O (exact in FP64), and
//            Q0 - 0.5, E_k = (Q_k - Q_0) / (4n) rounded to float.  Records go to a per-warp slab of shared memory.
//   phase B  8 lanes <-> tet (4 tets at a time, each group walks its own 8 tets of the tile): lane j takes table rows
//            j, j+8, ... in batches of 4 (indices first, then 4 gathers in flight, then accumulate).
// Lattice form of the sample position: the reference's 6-decimal weights are b_k = a_k/(4n) + eta_k with integer a_k,
// sum a_k = 4n, |eta_k| <= 5e-7, eps = sum b_k - 1 (checked on the host when the table is built).  Hence, exactly,
//     sum_k b_k P_k - O = Q_0 + sum_{k=1..3} a_k E_k + eps * O + tau,   |tau| <= 4 * 5e-7 * Rmax,  Rmax = max |Q_k|
// so the FP32 path needs 12 FFMA per sample and its inputs are small numbers (|Q| < ~20 instead of ~512), which shrinks
// the error bound -- and with it the fraction of samples that must be re-evaluated exactly -- by two orders of magnitude:
//     |q - r| <= (41 Rmax + 5) * 2^-24 ;  delta = (48 Rmax + 8) * 2^-24.
// floor() uses the 1.5*2^23 trick (no F2I), the voxel is widened with integer ops (no F2F): the XU pipe stays idle.
// One guard-band test per batch (largest |q - round(q)| of the lane's four samples against 1/2 - delta); only when it fires are the
// samples looked at one by one and the unsafe ones re-evaluated with the reference's FP64 sequence.
// Tets reaching outside the volume, slivers along one of its faces (phase A has the rule and the argument), very large tets
// (n > kSmemLevels) and tets wider than 100 voxels use the exact path.
// -------------------------------------------------------------------------------------------------
// BRICK (variant bit 2048): the same kernel gathering from the bricked, padded copy of the volume.  Origins are rounded down
// to brick corners, so idx = base + rx + 28 hx + 4 ry + (Sy-8) hy + 8 rz + (Sz-32) hz with hx = floor(rx/4), hy = floor(ry/2),
// hz = floor(rz/4), each from one more FFMA on the rounded value (dsc_tet_g32.cuh explains the trick).  A gather then touches
// 9.4 distinct 128 B lines instead of 16.8 (scripts/line_sim.py), and every tet inside the padded region -- including the
// layer of tets along the volume border -- takes the filtered path.
// PACK (variant bit 8192): the position arithmetic of a lane's four samples runs as two packed pairs (FFMA2 / FADD2), from a
// copy of the lattice table laid out in pairs of rows (i, i + 8).
// LEAN (variant bit 16384, part of the default): full batches run without row clamps and masks; only the last, partial batch of a
// tet carries them.  LEAN == 2 (bit 32768 as well): two full batches at a time -- eight gathers in flight per lane, one wait
// instead of two for a 64-sample tet.  Sums-only instantiations (no per-tet outputs) take v / n^3 from phase A instead of
// dividing once per round, whatever LEAN is.
// PF (variant bit 65536): one lane asks the TMA unit to pull the voxel box of the warp's next tile into L2 (tile_prefetch_l2).
// per-tet record of the w2 kernel: 112 bytes = 28 words, so the four records a round reads (and the records eight lanes write) fall in distinct banks
struct alignas(16) W2Rec {
    float4 q0, e1, e2, e3; // (Q0 - .5 xyz, thr) (E1 xyz, Ox) (E2 xyz, Oy) (E3 xyz, Oz)
    uint4 nd;              // node ids (exact fallback)
    int2 m;                // meta, (O_linear - fold)
    double v, rw;          // volume, v / n^3
    double pad;
};
static_assert(sizeof(W2Rec) == 112, "W2Rec layout");

template <typename VoxT, bool WANT_SQ, bool PER_TET, bool BRICK = false, bool SIMPLE = false, bool PACK = false, int LEAN = 0, bool PF = false>
__global__ void __launch_bounds__(256, W2_MIN_BLOCKS) tet_integrate_w2_kernel(const __grid_constant__ TetArgs a)
{
    constexpr int G = 8, U = 4;
    constexpr bool F32S = SIMPLE && sizeof(VoxT) == 4 && !VoxTraits<VoxT>::integral; // sums scaled by 2^-896 (widen_scaled)
    constexpr int kPairEntries = 9 * 2 * 8;    // levels n <= 5: 1+1+1+2+4 iterations of 32 rows, two row pairs per lane and iteration
    __shared__ float4 s_lat[PACK ? (2 * kPairEntries) : kSmemRows]; // (a1, a2, a3, eps) per table row; PACK: pairs, see below
    __shared__ double2 s_tabd[2 * kSmemRows];  // exact rows for the fallback
    // one record per tet of the warp's tile, written in phase A (lane <-> tet), read in phase B (8 lanes <-> tet): a single array, so a
    // round addresses all its fields from one base (eight separate arrays cost ~40 address instructions per round under the register cap)
    __shared__ W2Rec w_rec[8][32];
    constexpr bool RW = LEAN != 0 || !PER_TET; // sums only: v / n^3 once per tet in phase A (lanes in parallel) instead of a division per round
    __shared__ unsigned long long w_ns[8];     // samples processed by each warp (kept out of the registers of phase B)
    for (int i = threadIdx.x; i < 2 * kSmemRows; i += blockDim.x) s_tabd[i] = reinterpret_cast<const double2 *>(a.tet_bary)[i];
    if (threadIdx.x < 8) w_ns[threadIdx.x] = 0;
    if (PACK) {
        // entry e = (iteration slot * 2 + p) * 8 + g holds rows i = 32 it + 16 p + g and j = i + 8 of its level (clamped to the last
        // row) as (a1_i, a1_j, a2_i, a2_j), (a3_i, a3_j, eps_i, eps_j): each LDS.128 yields two register pairs ready for FFMA2
        for (int e = threadIdx.x; e < kPairEntries; e += blockDim.x) {
            const int slot = e >> 4, pp = (e >> 3) & 1, g = e & 7;
            int L = 0;
            while (L < kSmemLevels - 1 && g32_slot_off(L + 1) <= slot) L++;
            const int it = slot - g32_slot_off(L), n3 = (L + 1) * (L + 1) * (L + 1), off = (L * (L + 1) / 2) * (L * (L + 1) / 2);
            const int i = min(32 * it + 16 * pp + g, n3 - 1), j = min(32 * it + 16 * pp + g + 8, n3 - 1);
            const float4 ri = a.tet_lat[off + i], rj = a.tet_lat[off + j];
            s_lat[2 * e] = make_float4(ri.x, rj.x, ri.y, rj.y);
            s_lat[2 * e + 1] = make_float4(ri.z, rj.z, ri.w, rj.w);
        }
    } else {
        for (int i = threadIdx.x; i < kSmemRows; i += blockDim.x) s_lat[i] = a.tet_lat[i];
    }
    __syncthreads();

    const VoxT *__restrict__ vol = static_cast<const VoxT *>(BRICK ? a.brick : a.vol.data);
    const int X = a.vol.X, Y = a.vol.Y, Z = a.vol.Z;
    const int XYi = X * Y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lane_g = lane & (G - 1), grp = lane >> 3;
    const unsigned n_work = a.tet_end - a.tet_begin;
    const unsigned n_tiles = (n_work + 31) / 32;
    const unsigned total_warps = gridDim.x * 8;
    const float MAGIC = 12582912.0f; // 1.5 * 2^23
    const unsigned bSy = a.brick_sy, bSz = a.brick_sz, bKy = bSy - 8u, bKz = bSz - 32u;
    // bits of 1.5*2^23, 1.5*2^24, 1.5*2^25 times the coefficients they enter the index with
    const unsigned fold = BRICK ? 0x4B400000u * 13u + 0x4C400000u * (28u + bKz) + 0x4BC00000u * bKy : 0x4B400000u * (unsigned)(XYi + X + 1);
    // exact sample voxel as an element index of the volume this kernel reads; -1: outside (linear form only)
    auto exact_index = [&](int ix, int iy, int iz) -> int {
        const bool inb = (unsigned)ix < (unsigned)X && (unsigned)iy < (unsigned)Y && (unsigned)iz < (unsigned)Z;
        if (BRICK) return inb ? (int)brick_index(ix, iy, iz, bSy, bSz) : 0; // element 0 is padding: reads 0
        return inb ? iz * XYi + iy * X + ix : -1;
    };

    double acc_tot = 0, acc_sq = 0, acc_vol = 0; // per-label sums of label == lane_g + 1

    for (unsigned tile = blockIdx.x * 8 + warp; tile < n_tiles; tile += total_warps) {
        const unsigned base_t = a.tet_begin + tile * 32;
        const int nb = (int)min(32u, a.tet_end - base_t);
        __syncwarp();
        // ---------------- phase A: lane <-> tet ----------------
        // Records are stored in a permuted order: tets with at most 32 samples first, then the larger ones, inactive tets
        // last -- the four groups of a phase-B round then mostly run the same number of batches (a 27-sample tet next to a
        // 64-sample one would idle for a batch) and rounds past the last active tet are skipped.
        int n_act = 0;
        {
            int meta = 0, olin = 0;
            bool active = false, big = false, fast = false;
            int n3_lane = 0;
            float4 rq0 = make_float4(0, 0, 0, 0), re1 = rq0, re2 = rq0, re3 = rq0;
            double rv = 0;
            uint4 rnd = make_uint4(0, 0, 0, 0);
            if (lane < nb) {
                const unsigned t = base_t + lane;
                const int label = __ldg(a.mesh.tet_label + t);
                const uint4 nd = __ldg(a.mesh.tet_nodes + t); // with the label, not after it: one dependent round trip less per tile
                active = tet_active(label, a.include_bound);
                meta = (label & 0xff) | (lane << 16);
                if (active) {
                    const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                    const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
                    const double v = tet_volume(p0, p1, p2, p3);
                    const int L = tet_level(v);
                    n3_lane = (L + 1) * (L + 1) * (L + 1);
                    big = n3_lane > G * U;
                    const double lox = fmin(fmin(p0.x, p1.x), fmin(p2.x, p3.x)), hix = fmax(fmax(p0.x, p1.x), fmax(p2.x, p3.x));
                    const double loy = fmin(fmin(p0.y, p1.y), fmin(p2.y, p3.y)), hiy = fmax(fmax(p0.y, p1.y), fmax(p2.y, p3.y));
                    const double loz = fmin(fmin(p0.z, p1.z), fmin(p2.z, p3.z)), hiz = fmax(fmax(p0.z, p1.z), fmax(p2.z, p3.z));
                    // Inside the volume: no sample can leave [0, dim), so no bounds checks and no negative coordinates (trunc == floor).
                    // With a margin that is plain.  Without one -- the first layer of tets of a DSC mesh has nodes exactly on the faces of
                    // the volume -- it still holds when the tet is not a sliver along that axis: every weight of every row is at least
                    // b_min = 1/(4n) - 5e-7 >= 0.05 (a_k >= 1 for n <= 5, checked when the table is built) and the weights sum to 1 + eps,
                    // |eps| <= 2.1e-6, so  lo (1 + eps) + b_min (hi - lo) <= sum_k b_k p_k <= hi (1 + eps) - b_min (hi - lo):  with
                    // -1e-7 <= lo, hi <= dim + 1e-7 and hi - lo >= 1e-4 dim the sample coordinate lies in (0, dim).
                    // (BRICK: inside the padded copy with a margin of two voxels; the padding stands in for the bounds test)
                    auto axis_in = [](double lo, double hi, int dim) { return lo >= -1.0e-7 && hi <= (double)dim + 1.0e-7 && hi - lo >= 1.0e-4 * (double)dim; };
                    const bool interior = BRICK ? (lox >= -6.0 && loy >= -6.0 && loz >= -6.0 && hix < (double)X + 6.0 && hiy < (double)Y + 6.0 && hiz < (double)Z + 6.0)
                                                : ((lox >= 1.5 && loy >= 1.5 && loz >= 1.5 && hix < (double)X - 1.5 && hiy < (double)Y - 1.5 && hiz < (double)Z - 1.5) ||
                                                   (axis_in(lox, hix, X) && axis_in(loy, hiy, Y) && axis_in(loz, hiz, Z)));
                    double ox = floor(lox), oy = floor(loy), oz = floor(loz);
                    if (BRICK && interior) { // round the origin down to a brick corner (the padding is a multiple of the brick)
                        ox = (double)((((int)ox + kBrickPad) & ~3) - kBrickPad);
                        oy = (double)((((int)oy + kBrickPad) & ~1) - kBrickPad);
                        oz = (double)((((int)oz + kBrickPad) & ~3) - kBrickPad);
                    }
                    const double rmax = fmax(fmax(hix - ox, hiy - oy), hiz - oz);
                    fast = interior && L < kSmemLevels && rmax < 100.0;
                    meta |= (L << 8) | (1 << 12);
                    if (fast) {
                        const double inv4n = 1.0 / (double)(4 * (L + 1));
                        rq0 = make_float4((float)((p0.x - ox) - 0.5), (float)((p0.y - oy) - 0.5), (float)((p0.z - oz) - 0.5),
                                          0.5f - (48.0f * (float)rmax + 8.0f) * 5.9604644775390625e-08f);
                        re1 = make_float4((float)((p1.x - p0.x) * inv4n), (float)((p1.y - p0.y) * inv4n), (float)((p1.z - p0.z) * inv4n), (float)ox);
                        re2 = make_float4((float)((p2.x - p0.x) * inv4n), (float)((p2.y - p0.y) * inv4n), (float)((p2.z - p0.z) * inv4n), (float)oy);
                        re3 = make_float4((float)((p3.x - p0.x) * inv4n), (float)((p3.y - p0.y) * inv4n), (float)((p3.z - p0.z) * inv4n), (float)oz);
                        olin = BRICK ? (int)(brick_index((int)ox, (int)oy, (int)oz, bSy, bSz) - fold) : (int)((unsigned)((int)oz * XYi + (int)oy * X + (int)ox) - fold);
                        meta |= 1 << 13;
                    }
                    rv = v;
                    rnd = nd;
                } else {
                    if (a.tet_total) a.tet_total[t] = -1.0;
                    if (a.tet_vol) a.tet_vol[t] = -1.0;
                    if (a.tet_mean) a.tet_mean[t] = -1.0;
                }
            }
            const unsigned m_small = __ballot_sync(0xffffffffu, active && !big);
            const unsigned m_big = __ballot_sync(0xffffffffu, active && big);
            const unsigned lt = (1u << lane) - 1u;
            n_act = __popc(m_small) + __popc(m_big);
            {
                const unsigned tile_ns = __reduce_add_sync(0xffffffffu, (unsigned)n3_lane);
                if (lane == 0) w_ns[warp] += tile_ns;
            }
            const int slot = active ? (big ? __popc(m_small) + __popc(m_big & lt) : __popc(m_small & lt))
                                    : n_act + __popc(~(m_small | m_big) & lt);
            w_rec[warp][slot].m = make_int2(meta, olin);
            if (active) {
                w_rec[warp][slot].v = rv;
                if (RW) w_rec[warp][slot].rw = rv / (double)n3_lane;
                w_rec[warp][slot].nd = rnd;
                if (fast) { w_rec[warp][slot].q0 = rq0; w_rec[warp][slot].e1 = re1; w_rec[warp][slot].e2 = re2; w_rec[warp][slot].e3 = re3; }
            }
        }
        __syncwarp();
        // ---------------- phase B: 8 lanes <-> tet; group grp walks tets grp, grp+4, ... of the tile ----------------
        for (int r = 0; r * 4 < n_act; r++) {
            const int j = r * 4 + grp;
            const int2 mm = w_rec[warp][j].m;
            const int meta = mm.x;
            const bool active = (meta & (1 << 12)) != 0;
            const int label = meta & 0xff;
            const int L = (meta >> 8) & 0xf;
            const int n3 = active ? (L + 1) * (L + 1) * (L + 1) : 0;
            const int off = (L * (L + 1) / 2) * (L * (L + 1) / 2);
            double s = 0, sq = 0;
            unsigned long long isum = 0, isq = 0;
            auto add_voxel = [&](VoxT val) {
                if constexpr (VoxTraits<VoxT>::integral) {
                    const unsigned q = (unsigned)val;
                    isum += q;
                    if (WANT_SQ) isq += (unsigned long long)q * q;
                } else {
                    double I;
                    if constexpr (F32S) I = widen_scaled((float)val);
                    else if constexpr (sizeof(VoxT) == 4) I = widen_exact((float)val);
                    else I = (double)val;
                    s += I;
                    if (WANT_SQ) {
                        const double Iu = F32S ? I * 0x1p+896 : I; // exact
                        sq = fma(Iu, Iu, sq); // second moment: tolerance quantity, fused on purpose
                    }
                }
            };
            if (meta & (1 << 13)) {
                const float4 q0 = w_rec[warp][j].q0, e1 = w_rec[warp][j].e1, e2 = w_rec[warp][j].e2, e3 = w_rec[warp][j].e3;
                const float thr = q0.w;
                const unsigned obase = (unsigned)mm.y;
                const float4 *__restrict__ tl = s_lat + off;
                const ulonglong2 *__restrict__ tl2 = reinterpret_cast<const ulonglong2 *>(s_lat) + (size_t)g32_slot_off(L) * 32 + 2 * lane_g;
                const int last = n3 - 1;
                // one batch of 4 samples per lane; TAIL: the batch may reach past the last row of the level (rows are clamped and masked)
                auto batch = [&](auto tail_c, auto nb_c, int base, int it) {
                    constexpr bool TAIL = decltype(tail_c)::value;
                    constexpr int U = 4 * decltype(nb_c)::value; // samples per lane in this call (shadows the kernel's U)
                    int idx[U];
                    unsigned unsafe = 0;
                    float emax[U]; // largest |q - round(q)| of each sample; one test per batch on the common path
                    if constexpr (PACK) {
                        const f32x2 M2 = pk2(MAGIC, MAGIC), NM2 = pk2(-MAGIC, -MAGIC), NEG1 = pk2(-1.0f, -1.0f);
#pragma unroll
                        for (int pp = 0; pp < 2; pp++) { // rows (i, i + 8) in the two halves of a packed register
                            const ulonglong2 A = tl2[(it * 2 + pp) * 16], B = tl2[(it * 2 + pp) * 16 + 1];
                            const int i0 = base + 16 * pp;
                            const bool ok0 = !TAIL || i0 <= last, ok1 = !TAIL || i0 + 8 <= last;
                            const f32x2 a1 = A.x, a2 = A.y, a3 = B.x, ep = B.y;
                            f32x2 qx = fma2(ep, pk2(e1.w, e1.w), pk2(q0.x, q0.x)), qy = fma2(ep, pk2(e2.w, e2.w), pk2(q0.y, q0.y));
                            f32x2 qz = fma2(ep, pk2(e3.w, e3.w), pk2(q0.z, q0.z));
                            qx = fma2(a1, pk2(e1.x, e1.x), qx); qy = fma2(a1, pk2(e1.y, e1.y), qy); qz = fma2(a1, pk2(e1.z, e1.z), qz);
                            qx = fma2(a2, pk2(e2.x, e2.x), qx); qy = fma2(a2, pk2(e2.y, e2.y), qy); qz = fma2(a2, pk2(e2.z, e2.z), qz);
                            qx = fma2(a3, pk2(e3.x, e3.x), qx); qy = fma2(a3, pk2(e3.y, e3.y), qy); qz = fma2(a3, pk2(e3.z, e3.z), qz);
                            const f32x2 tx2 = add2(qx, M2), ty2 = add2(qy, M2), tz2 = add2(qz, M2);
                            const f32x2 ex2 = fma2(add2(tx2, NM2), NEG1, qx), ey2 = fma2(add2(ty2, NM2), NEG1, qy), ez2 = fma2(add2(tz2, NM2), NEG1, qz);
                            float tx[2], ty[2], tz[2], ex[2], ey[2], ez[2];
                            upk2(tx2, tx[0], tx[1]); upk2(ty2, ty[0], ty[1]); upk2(tz2, tz[0], tz[1]);
                            upk2(ex2, ex[0], ex[1]); upk2(ey2, ey[0], ey[1]); upk2(ez2, ez[0], ez[1]);
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                const bool ok = k ? ok1 : ok0;
                                const int lin = (int)(__float_as_uint(tz[k]) * (unsigned)XYi + __float_as_uint(ty[k]) * (unsigned)X + __float_as_uint(tx[k]) + obase);
                                idx[2 * pp + k] = ok ? lin : -1;
                                emax[2 * pp + k] = fmaxf(fmaxf(fabsf(ex[k]), fabsf(ey[k])), fabsf(ez[k]));
                            }
                        }
                    } else {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = base + u * G;
                        const bool ok = !TAIL || i <= last;
                        const float4 c = tl[TAIL ? min(i, last) : i];
                        const float qx = fmaf(c.z, e3.x, fmaf(c.y, e2.x, fmaf(c.x, e1.x, fmaf(c.w, e1.w, q0.x))));
                        const float qy = fmaf(c.z, e3.y, fmaf(c.y, e2.y, fmaf(c.x, e1.y, fmaf(c.w, e2.w, q0.y))));
                        const float qz = fmaf(c.z, e3.z, fmaf(c.y, e2.z, fmaf(c.x, e1.z, fmaf(c.w, e3.w, q0.z))));
                        const float tx = qx + MAGIC, ty = qy + MAGIC, tz = qz + MAGIC;
                        const float ex = qx - (tx - MAGIC), ey = qy - (ty - MAGIC), ez = qz - (tz - MAGIC);
                        int lin;
                        if (BRICK) {
                            const float hx = fmaf(tx, 0.99999988079071044921875f, 37748736.0f), hy = fmaf(ty, 0.999999940395355224609375f, MAGIC);
                            const float hz = fmaf(tz, 0.99999988079071044921875f, 37748736.0f);
                            lin = (int)(__float_as_uint(tx) + 28u * __float_as_uint(hx) + 4u * __float_as_uint(ty) + bKy * __float_as_uint(hy) +
                                        8u * __float_as_uint(tz) + bKz * __float_as_uint(hz) + obase);
                        } else {
                            lin = (int)(__float_as_uint(tz) * (unsigned)XYi + __float_as_uint(ty) * (unsigned)X + __float_as_uint(tx) + obase);
                        }
                        idx[u] = ok ? lin : -1;
                        emax[u] = fmaxf(fmaxf(fabsf(ex), fabsf(ey)), fabsf(ez)); // rows past the last one repeat it: harmless in the batch-wide test
                    }
                    }
                    {
                        float eall = emax[0];
#pragma unroll
                        for (int u = 1; u < U; u++) eall = fmaxf(eall, emax[u]);
                        if (eall >= thr) {
#pragma unroll
                            for (int u = 0; u < U; u++)
                                if (idx[u] >= 0 && emax[u] >= thr) unsafe |= 1u << u;
                        }
                    }
                    if (unsafe) { // a coordinate within delta of a voxel boundary: the reference's FP64 arithmetic decides
                        const uint4 nd = w_rec[warp][j].nd;
                        const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                        const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
#pragma unroll
                        for (int u = 0; u < U; u++)
                            if (unsafe & (1u << u)) {
                                const int i = base + u * G;
                                const double2 b01 = s_tabd[2 * (off + i)], b23 = s_tabd[2 * (off + i) + 1];
                                const int ix = (int)(((p0.x * b01.x + p1.x * b01.y) + p2.x * b23.x) + p3.x * b23.y);
                                const int iy = (int)(((p0.y * b01.x + p1.y * b01.y) + p2.y * b23.x) + p3.y * b23.y);
                                const int iz = (int)(((p0.z * b01.x + p1.z * b01.y) + p2.z * b23.x) + p3.z * b23.y);
                                idx[u] = BRICK ? exact_index(ix, iy, iz) : iz * XYi + iy * X + ix; // linear form: interior tet, always in range
                            }
                    }
                    VoxT val[U];
#pragma unroll
                    for (int u = 0; u < U; u++) val[u] = (!TAIL || idx[u] >= 0) ? __ldg(vol + idx[u]) : (VoxT)0;
#pragma unroll
                    for (int u = 0; u < U; u++) add_voxel(val[u]);
                };
                using N1 = std::integral_constant<int, 1>;
                using N2 = std::integral_constant<int, 2>;
                if (LEAN) { // full batches first: no clamps, no masks
                    int base = lane_g, it = 0;
                    if (LEAN == 2 && !PACK)
                        for (; base - lane_g + 2 * G * U <= n3; base += 2 * G * U, it += 2) batch(std::false_type{}, N2{}, base, it);
                    for (; base - lane_g + G * U <= n3; base += G * U, it++) batch(std::false_type{}, N1{}, base, it);
                    for (; base < n3; base += G * U, it++) batch(std::true_type{}, N1{}, base, it);
                } else {
                    for (int base = lane_g, it = 0; base < n3; base += G * U, it++) batch(std::true_type{}, N1{}, base, it);
                }
            } else if (active) {
                // exact path (get_coord, DEMO/tet_dis_coord.hpp:27; truncation image3d.cpp:366; bounds image3d.cpp:480-491)
                const uint4 nd = w_rec[warp][j].nd;
                const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
                const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(a.tet_bary) + 2 * (size_t)off;
                for (int base = lane_g; base < n3; base += G * U) {
                    int idx[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = base + u * G;
                        idx[u] = -1;
                        if (i < n3) {
                            const double2 b01 = __ldg(tb + 2 * i), b23 = __ldg(tb + 2 * i + 1);
                            const int ix = (int)(((p0.x * b01.x + p1.x * b01.y) + p2.x * b23.x) + p3.x * b23.y);
                            const int iy = (int)(((p0.y * b01.x + p1.y * b01.y) + p2.y * b23.x) + p3.y * b23.y);
                            const int iz = (int)(((p0.z * b01.x + p1.z * b01.y) + p2.z * b23.x) + p3.z * b23.y);
                            idx[u] = exact_index(ix, iy, iz);
                        }
                    }
                    VoxT val[U];
#pragma unroll
                    for (int u = 0; u < U; u++) val[u] = idx[u] >= 0 ? __ldg(vol + idx[u]) : (VoxT)0;
#pragma unroll
                    for (int u = 0; u < U; u++) add_voxel(val[u]);
                }
            }
            if constexpr (VoxTraits<VoxT>::integral) {
                const double dn = VoxTraits<VoxT>::denom();
                s = (double)isum / dn;
                if (WANT_SQ) sq = (double)isq / (dn * dn);
            }
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) {
                s += shfl_xor_d(s, m);
                if (WANT_SQ) sq += shfl_xor_d(sq, m);
            }
            if (F32S) s = s * 0x1p+896; // exact: back to the true sum
            if (active) {
                const double v = w_rec[warp][j].v;
                // total = total * v / a.size()  (image3d.cpp:371) when per-tet values go out; the per-label sums alone are 1e-6 quantities
                const double tot = (RW && !PER_TET) ? s * w_rec[warp][j].rw : s * v / (double)n3;
                const double tsq = WANT_SQ ? ((RW && !PER_TET) ? sq * w_rec[warp][j].rw : sq * v / (double)n3) : 0.0;
                if (lane_g == 0) {
                    if (PER_TET) {
                        const unsigned t = base_t + ((meta >> 16) & 31);
                        if (a.tet_total) a.tet_total[t] = tot;
                        if (a.tet_vol) a.tet_vol[t] = v;
                        if (a.tet_mean) a.tet_mean[t] = tot / v;
                    }
                }
                if (label >= 1 && label <= a.nb_phase && lane_g == label - 1) {
                    acc_tot += tot;
                    acc_sq += tsq;
                    acc_vol += v;
                }
            }
        }
    }
#pragma unroll
    for (int m = 16; m >= G; m >>= 1) {
        acc_tot += shfl_xor_d(acc_tot, m);
        acc_sq += shfl_xor_d(acc_sq, m);
        acc_vol += shfl_xor_d(acc_vol, m);
    }
    __shared__ double sm[8][3 * kMaxPhase];
    if (lane < kMaxPhase) {
        sm[warp][lane] = acc_tot;
        sm[warp][kMaxPhase + lane] = acc_sq;
        sm[warp][2 * kMaxPhase + lane] = acc_vol;
    }
    __syncthreads();
    if (threadIdx.x < 3 * kMaxPhase) {
        double r = 0;
        for (int wdx = 0; wdx < 8; wdx++) r += sm[wdx][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 3 * kMaxPhase + threadIdx.x] = r;
    }
    if (lane == 0 && w_ns[warp]) atomicAdd(a.sample_count, w_ns[warp]);
}

// Adds the per-CTA partial rows and derives the means: out[0..P) total, [P..2P) sumsq, [2P..3P) volume (P = nb_phase),
// mean[k] = total/volume.  One warp per column group: lane l adds rows l, l+32, ... in order, then a fixed xor tree --
// a deterministic order that depends only on the number of rows (a single thread per column took 70 us for 444 rows).
__global__ void __launch_bounds__(32 * 3 * kMaxPhase) tet_finalize_kernel(const double *__restrict__ partials, int n_rows, int nb_phase, double *sums, double *mean)
{
    const int col = threadIdx.x >> 5, lane = threadIdx.x & 31; // 24 warps, one per column
    double r = 0;
    for (int i = lane; i < n_rows; i += 32) r += partials[(size_t)i * 3 * kMaxPhase + col];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) r += shfl_xor_d(r, m);
    const int which = col / kMaxPhase, ph = col % kMaxPhase;
    __shared__ double tot[kMaxPhase], vol[kMaxPhase];
    if (lane == 0) {
        if (ph < nb_phase) sums[which * nb_phase + ph] = r;
        if (which == 0) tot[ph] = r;
        if (which == 2) vol[ph] = r;
    }
    __syncthreads();
    if (mean && threadIdx.x < nb_phase) mean[threadIdx.x] = tot[threadIdx.x] / vol[threadIdx.x];
}

__global__ void means_from_sums_kernel(const double *sums, int nb_phase, double *mean)
{
    const int k = threadIdx.x;
    if (k < nb_phase) mean[k] = sums[k] / sums[2 * nb_phase + k];
}

// Neighbour-area term of get_energy_tetrahedron (DEMO/segment_function.cpp:629-640) added to the per-candidate variation
// sums that the tet pass left in `energy`: one thread per (tet, candidate), faces in get_faces(t) order.
__global__ void relabel_area_kernel(MeshView m, const unsigned *__restrict__ tet_face_nodes, const unsigned *__restrict__ tet_face_tets, int nb_phase,
                                    double alpha, double *energy)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)m.n_tets * nb_phase) return;
    const unsigned t = (unsigned)(i / nb_phase);
    const int k = (int)(i % nb_phase);
    const int old_label = m.tet_label[t];
    if (old_label <= 0) { energy[i] = -1.0; return; }
    double e = energy[i];
    for (int f = 0; f < 4; f++) {
        const unsigned ta = tet_face_tets[((size_t)t * 4 + f) * 2], tb = tet_face_tets[((size_t)t * 4 + f) * 2 + 1];
        if (ta == kNoTet || tb == kNoTet) continue;
        int label0 = m.tet_label[ta];
        const int label1 = m.tet_label[tb];
        label0 = label0 == old_label ? label1 : label0;
        if (label0 != k + 1) {
            const unsigned *fn = tet_face_nodes + ((size_t)t * 4 + f) * 3;
            e += tri_area(load_pos(m.pos4, fn[0]), load_pos(m.pos4, fn[1]), load_pos(m.pos4, fn[2])) * alpha;
        }
    }
    energy[i] = e;
}

// Bit-exactness witness: one thread per tet walks its sample points in table order and hashes the
// truncated voxel coordinates (order-sensitive FNV-1a), exactly what oracle/ref_driver.cpp records
// with the reference's own table and macro.
__global__ void tet_sample_hash_kernel(MeshView mesh, const double *__restrict__ tet_bary, unsigned long long *hash, int *nsamp)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= mesh.n_tets) return;
    const uint4 nd = __ldg(mesh.tet_nodes + t);
    const V3 p0 = load_pos(mesh.pos4, nd.x), p1 = load_pos(mesh.pos4, nd.y), p2 = load_pos(mesh.pos4, nd.z), p3 = load_pos(mesh.pos4, nd.w);
    const double v = tet_volume(p0, p1, p2, p3);
    const int L = tet_level(v);
    const int n3 = (L + 1) * (L + 1) * (L + 1);
    const int off = (L * (L + 1) / 2) * (L * (L + 1) / 2);
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < n3; i++) {
        const double *b = tet_bary + 4 * (size_t)(off + i);
        const double px = ((p0.x * b[0] + p1.x * b[1]) + p2.x * b[2]) + p3.x * b[3];
        const double py = ((p0.y * b[0] + p1.y * b[1]) + p2.y * b[2]) + p3.y * b[3];
        const double pz = ((p0.z * b[0] + p1.z * b[1]) + p2.z * b[2]) + p3.z * b[3];
        const int c[3] = {(int)px, (int)py, (int)pz};
#pragma unroll
        for (int k = 0; k < 3; k++) { h ^= (unsigned long long)(unsigned)c[k]; h *= 1099511628211ull; }
    }
    hash[t] = h;
    nsamp[t] = n3;
}

// is_point_inside (DEMO/segment_function.cpp:2576-2580) on the 4-point Util::barycentric_coords
// (is_mesh/util.h:349-364): all four normalised signed volumes >= 0.
__global__ void points_in_tets_kernel(MeshView mesh, unsigned n, const unsigned *tet_idx, const double *pts, unsigned char *inside)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 nd = __ldg(mesh.tet_nodes + tet_idx[i]);
    const V3 a = load_pos(mesh.pos4, nd.x), b = load_pos(mesh.pos4, nd.y), c = load_pos(mesh.pos4, nd.z), d = load_pos(mesh.pos4, nd.w);
    const V3 p = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    double c0 = signed_volume(p, b, c, d), c1 = signed_volume(a, p, c, d), c2 = signed_volume(a, b, p, d), c3 = signed_volume(a, b, c, p);
    const double s = ((c0 + c1) + c2) + c3;
    c0 /= s; c1 /= s; c2 /= s; c3 /= s;
    inside[i] = (c0 >= 0 && c1 >= 0 && c2 >= 0 && c3 >= 0) ? 1 : 0;
}

// =================================================================================================
// External force: segment_function::compute_external_force (DEMO/segment_function.cpp:1425-1473)
// =================================================================================================
struct FaceGeom {
    V3 p[3];
    V3 norm;
    double area, c0, c1;
    int gs;
};

DSC_D int first_not_in(uint4 tn, unsigned f0, unsigned f1, unsigned f2)
{   // (get_nodes(tid) - nids).front(), is_mesh/is_mesh.h:632-637
    if (tn.x != f0 && tn.x != f1 && tn.x != f2) return tn.x;
    if (tn.y != f0 && tn.y != f1 && tn.y != f2) return tn.y;
    if (tn.z != f0 && tn.z != f1 && tn.z != f2) return tn.z;
    return tn.w;
}

// DSC::get_normal(fid, tid): normal_direction over get_sorted_nodes(fid, tid)
// (src/DSC.h:3120-3124; is_mesh/is_mesh.h:632-637, 981-1000; is_mesh/util.h:366-376)
DSC_D V3 face_normal_wrt(const MeshView &m, unsigned f0, unsigned f1, unsigned f2, uint4 tn)
{
    const unsigned apex = first_not_in(tn, f0, f1, f2);
    const V3 A = load_pos(m.pos4, f0), B = load_pos(m.pos4, f1), C = load_pos(m.pos4, f2), Q = load_pos(m.pos4, apex);
    const V3 x = sub(Q, A), y = sub(B, A), z = sub(C, A);
    const bool cw = dot(x, cross(y, z)) > 0.;
    const V3 a = cw ? B : A, b = cw ? A : B;
    return normalize(cross(sub(b, a), sub(C, a)));
}

DSC_D V3 tet_barycenter(const MeshView &m, uint4 tn)
{   // (a + b + c + d)*0.25 in get_nodes(t) order (src/DSC.h:3254-3258, is_mesh/util.h:100-103)
    const V3 a = load_pos(m.pos4, tn.x), b = load_pos(m.pos4, tn.y), c = load_pos(m.pos4, tn.z), d = load_pos(m.pos4, tn.w);
    return mul(add(add(add(a, b), c), d), 0.25);
}

// the part of face_setup that does not depend on the phase means: nodes, positions, normal, area, Gauss set, the two labels
DSC_D bool face_geometry(const MeshView &m, unsigned f, FaceGeom &g, unsigned n[3], int &l0, int &l1)
{
    const unsigned t0 = __ldg(m.face_tets + 2 * (size_t)f), t1 = __ldg(m.face_tets + 2 * (size_t)f + 1);
    if (t0 == kNoTet || t1 == kNoTet) return false; // the reference would read tets[1] out of range here
    n[0] = __ldg(m.face_nodes + 3 * (size_t)f);
    n[1] = __ldg(m.face_nodes + 3 * (size_t)f + 1);
    n[2] = __ldg(m.face_nodes + 3 * (size_t)f + 2);
    g.p[0] = load_pos(m.pos4, n[0]); g.p[1] = load_pos(m.pos4, n[1]); g.p[2] = load_pos(m.pos4, n[2]);
    l0 = __ldg(m.tet_label + t0);
    l1 = __ldg(m.tet_label + t1);
    g.c0 = g.c1 = 0.0;
    const uint4 tn0 = __ldg(m.tet_nodes + t0), tn1 = __ldg(m.tet_nodes + t1);
    // get_normal(fid): sorted w.r.t. the co-boundary tet with the highest label, first wins (is_mesh.h:639-663)
    V3 N = face_normal_wrt(m, n[0], n[1], n[2], l1 > l0 ? tn1 : tn0);
    const V3 l01 = sub(tet_barycenter(m, tn1), tet_barycenter(m, tn0));
    N = mul(N, dot(N, l01));
    g.norm = divs(N, length(N)); // Vec3d::normalize
    g.area = tri_area(g.p[0], g.p[1], g.p[2]);
    g.gs = gauss_set(g.area);
    return true;
}

// phase_intensity (DEMO/segment_function.h:149-156): -0.3 for the label-0 side
DSC_D double phase_intensity(int label, int nb_phase, const double *__restrict__ mean) { return (label >= 1 && label <= nb_phase) ? mean[label - 1] : -0.3; }

DSC_D bool face_setup(const MeshView &m, unsigned f, int nb_phase, const double *__restrict__ mean, FaceGeom &g, unsigned n[3])
{
    int l0, l1;
    if (!face_geometry(m, f, g, n, l0, l1)) return false;
    g.c0 = phase_intensity(l0, nb_phase, mean);
    g.c1 = phase_intensity(l1, nb_phase, mean);
    return true;
}

struct ForceArgs {
    VolumeView vol;
    MeshView mesh;
    int nb_phase;
    const double *mean;
    double *forces;
    unsigned begin, end; // node range (gather) or face range (atomic)
    unsigned long long *gauss_count;
};

// Gather form: one thread per node; walks the node's incident interface faces in ascending face
// order and, inside a face, the Gauss points in table order -- the order in which the reference
// executes forces[verts[ii]] += f*c.coord[ii] (:1459-1468) -- so the result is bit-identical to the
// reference's, needs no atomics and does not depend on the number of GPUs.
template <typename VoxT> __global__ void __launch_bounds__(128) force_gather_kernel(const ForceArgs a)
{
    const unsigned v = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    double fx = 0, fy = 0, fz = 0;
    const unsigned e0 = __ldg(a.mesh.node_off + v), e1 = __ldg(a.mesh.node_off + v + 1);
    unsigned long long ng = 0;
    for (unsigned e = e0; e < e1; e++) {
        const unsigned inc = __ldg(a.mesh.node_inc + e);
        const unsigned f = inc >> 2, corner = inc & 3u;
        FaceGeom g;
        unsigned n[3];
        if (!face_setup(a.mesh, f, a.nb_phase, a.mean, g, n)) continue;
        for (int q = c_gauss_off[g.gs]; q < c_gauss_off[g.gs + 1]; q++) {
            const double w = c_gauss[4 * q], b0 = c_gauss[4 * q + 1], b1 = c_gauss[4 * q + 2], b2 = c_gauss[4 * q + 3];
            const V3 p = add(add(mul(g.p[0], b0), mul(g.p[1], b1)), mul(g.p[2], b2)); // get_coord_tri
            const double I = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
            const V3 fv = mul(g.norm, (2 * I - g.c0 - g.c1) * (g.c0 - g.c1) * w * g.area);
            const double bc = corner == 0 ? b0 : (corner == 1 ? b1 : b2);
            fx += fv.x * bc; fy += fv.y * bc; fz += fv.z * bc;
            if (corner == 0) ng++;
        }
    }
    a.forces[3 * (size_t)v] = fx;
    a.forces[3 * (size_t)v + 1] = fy;
    a.forces[3 * (size_t)v + 2] = fz;
}

// Warp-per-node gather (default): same result as force_gather_kernel, bit for bit, with the work of a node spread over
// the warp.  Lane l evaluates incident face l of the node (normal, area, Gauss set, trilinear samples) and stores its
// per-Gauss-point addends fv*coord to shared memory; lanes 0..2 then add the x/y/z components in the reference's order
// (faces ascending, Gauss points in table order).  Only nodes that have interface faces are visited (active list built
// with the CSR on the host); the force array is zeroed beforehand.
struct ForceWarpArgs {
    VolumeView vol;
    MeshView mesh;
    int nb_phase;
    const double *mean;
    double *forces;
    const unsigned *active_nodes;
    unsigned begin, end; // range in the active list
    unsigned long long *gauss_count;
};

template <typename VoxT> __global__ void __launch_bounds__(128) force_warp_kernel(const ForceWarpArgs a)
{
    __shared__ double slab[4][32][12][3];
    __shared__ int ng_s[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned item = a.begin + blockIdx.x * 4 + warp;
    if (item >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    const unsigned v = __ldg(a.active_nodes + item);
    const unsigned e0 = __ldg(a.mesh.node_off + v), e1 = __ldg(a.mesh.node_off + v + 1);
    double acc = 0; // lanes 0..2: x, y, z of the node's force
    unsigned long long ngc = 0;
    for (unsigned eb = e0; eb < e1; eb += 32) {
        const unsigned e = eb + lane;
        int ng = 0;
        if (e < e1) {
            const unsigned inc = __ldg(a.mesh.node_inc + e);
            const unsigned f = inc >> 2, corner = inc & 3u;
            FaceGeom g;
            unsigned n[3];
            if (face_setup(a.mesh, f, a.nb_phase, a.mean, g, n)) {
                for (int q = c_gauss_off[g.gs]; q < c_gauss_off[g.gs + 1]; q++) {
                    const double w = c_gauss[4 * q], b0 = c_gauss[4 * q + 1], b1 = c_gauss[4 * q + 2], b2 = c_gauss[4 * q + 3];
                    const V3 p = add(add(mul(g.p[0], b0), mul(g.p[1], b1)), mul(g.p[2], b2)); // get_coord_tri
                    const double I = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
                    const V3 fv = mul(g.norm, (2 * I - g.c0 - g.c1) * (g.c0 - g.c1) * w * g.area);
                    const double bc = corner == 0 ? b0 : (corner == 1 ? b1 : b2);
                    slab[warp][lane][ng][0] = fv.x * bc;
                    slab[warp][lane][ng][1] = fv.y * bc;
                    slab[warp][lane][ng][2] = fv.z * bc;
                    ng++;
                }
                if (corner == 0) ngc += (unsigned long long)ng;
            }
        }
        ng_s[warp][lane] = ng;
        __syncwarp();
        if (lane < 3) {
            const int cnt = (int)min(32u, e1 - eb);
            for (int l = 0; l < cnt; l++) {
                const int m = ng_s[warp][l];
                for (int q = 0; q < m; q++) acc += slab[warp][l][q][lane];
            }
        }
        __syncwarp();
    }
    if (lane < 3) a.forces[3 * (size_t)v + lane] = acc;
}

// Two-kernel deterministic gather (default):
//   force_face_kernel  one thread per interface face: normal, area, Gauss set, and for every Gauss point the vector
//                      f = Norm * ((2I - c0 - c1)(c0 - c1) w area) exactly as the reference forms it (:1463-1465),
//                      stored face-major (12 slots of 3 doubles per face) together with the Gauss-set index;
//   force_node_kernel  one thread per node that has interface faces: walks its CSR list (faces ascending) and adds
//                      f * c.coord[corner] Gauss point by Gauss point -- the reference's addition order (:1466-1467),
//                      so _forces is bit-identical, without atomics and independent of the GPU count.
// All faces' dependent load chains run concurrently (one thread each), and the node pass only reads the compact f buffer.
struct ForceFaceArgs {
    VolumeView vol;
    MeshView mesh;
    int nb_phase;
    const double *mean;
    double *fbuf;        // [n_faces][12][3]
    signed char *fset;   // [n_faces] Gauss set index, -1 = face skipped
    unsigned begin, end; // face range
    unsigned long long *gauss_count;
};

// kFaceLanes threads per face: thread s evaluates Gauss points s, s+kFaceLanes, ... of the face (sets have at most 12
// points), so the dependent load chain of a face (ids -> labels/tet ids -> positions -> 8 voxel corners) is walked by a
// few threads in parallel instead of once per Gauss point in sequence.
constexpr int kFaceLanes = 4;

template <typename VoxT> __global__ void __launch_bounds__(128) force_face_kernel(const ForceFaceArgs a)
{
    const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned f = a.begin + gt / kFaceLanes;
    const int slot = (int)(gt % kFaceLanes);
    if (f >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    FaceGeom g;
    unsigned n[3];
    if (!face_setup(a.mesh, f, a.nb_phase, a.mean, g, n)) {
        if (slot == 0) a.fset[f] = -1;
        return;
    }
    const int q0 = c_gauss_off[g.gs], ng = c_gauss_off[g.gs + 1] - q0;
    if (slot == 0) a.fset[f] = (signed char)g.gs;
    for (int k = slot; k < ng; k += kFaceLanes) {
        const int q = q0 + k;
        const double w = c_gauss[4 * q], b0 = c_gauss[4 * q + 1], b1 = c_gauss[4 * q + 2], b2 = c_gauss[4 * q + 3];
        const V3 p = add(add(mul(g.p[0], b0), mul(g.p[1], b1)), mul(g.p[2], b2)); // get_coord_tri
        const double I = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
        const V3 fv = mul(g.norm, (2 * I - g.c0 - g.c1) * (g.c0 - g.c1) * w * g.area);
        double *out = a.fbuf + (size_t)f * 36 + 3 * k;
        out[0] = fv.x; out[1] = fv.y; out[2] = fv.z;
    }
}

// ---- the same pass split at the only place the phase means enter -------------------------------------------------------------
// f = Norm * ((2I - c0 - c1)(c0 - c1) w area): everything but c0, c1 is known before the means are -- the normal, the area,
// the Gauss set and, per Gauss point, the trilinear sample I, which is where the time goes.  force_geom_kernel computes those
// (it can run next to the tet pass that produces the means, on a second stream), force_node2_kernel forms the vectors with
// the reference's operation order and adds them like force_node_kernel: the forces are bit-identical to the reference's.
struct ForceGeomArgs {
    VolumeView vol;
    MeshView mesh;
    double *fgeo;        // [n_faces][4]: unit normal (tet0 -> tet1), area
    int *flab;           // [n_faces][2]: labels of the two tets
    double *fint;        // [n_faces][12]: trilinear sample per Gauss point
    signed char *fset;   // [n_faces] Gauss set index, -1 = face skipped
    unsigned begin, end; // face range
};

// (1, 2 or 4 lanes per face time the same: the kernel moves 115 MB of 32-byte sectors for its scattered trilinear reads at C4 -- 3.5 TB/s)
template <typename VoxT> __global__ void __launch_bounds__(128) force_geom_kernel(const ForceGeomArgs a)
{
    const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned f = a.begin + gt / kFaceLanes;
    const int slot = (int)(gt % kFaceLanes);
    if (f >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    FaceGeom g;
    unsigned n[3];
    int l0, l1;
    if (!face_geometry(a.mesh, f, g, n, l0, l1)) {
        if (slot == 0) a.fset[f] = -1;
        return;
    }
    const int q0 = c_gauss_off[g.gs], ng = c_gauss_off[g.gs + 1] - q0;
    if (slot == 0) {
        a.fset[f] = (signed char)g.gs;
        double *o = a.fgeo + 4 * (size_t)f;
        o[0] = g.norm.x; o[1] = g.norm.y; o[2] = g.norm.z; o[3] = g.area;
        a.flab[2 * (size_t)f] = l0;
        a.flab[2 * (size_t)f + 1] = l1;
    }
    for (int k = slot; k < ng; k += kFaceLanes) {
        const int q = q0 + k;
        const double b0 = c_gauss[4 * q + 1], b1 = c_gauss[4 * q + 2], b2 = c_gauss[4 * q + 3];
        const V3 p = add(add(mul(g.p[0], b0), mul(g.p[1], b1)), mul(g.p[2], b2)); // get_coord_tri
        a.fint[12 * (size_t)f + k] = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
    }
}

struct ForceNode2Args {
    MeshView mesh;
    const double *fgeo;
    const int *flab;
    const double *fint;
    const signed char *fset;
    int nb_phase;
    const double *mean;
    double *forces;
    unsigned begin, end;          // node range handled by this context
    const unsigned *active_nodes; // nodes that have interface faces, ascending (built with the CSR)
    const unsigned *n_active;     // device-resident length of that list
};

// The same sums with one lane per component (x, y, z; the fourth lane of a quad idles): a lane walks the node's faces in ascending order and
// their Gauss points in table order and adds ITS component of every addend f * coord -- the reference's order of additions
// (DEMO/segment_function.cpp:1466-1467), no shared memory, no cross-lane traffic, every node of the mesh resident at once.
__global__ void __launch_bounds__(128) force_node3_kernel(const ForceNode2Args a)
{
    const int c = threadIdx.x & 3;
    const unsigned n_active = __ldg(a.n_active);
    for (unsigned item = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; item < n_active; item += (gridDim.x * blockDim.x) >> 2) {
        const unsigned v = __ldg(a.active_nodes + item);
        if (c == 3 || v < a.begin || v >= a.end) continue;
        const unsigned e0 = __ldg(a.mesh.node_off + v), e1 = __ldg(a.mesh.node_off + v + 1);
        double acc = 0;
#pragma unroll 2
        for (unsigned e = e0; e < e1; e++) {
            const unsigned inc = __ldg(a.mesh.node_inc + e);
            const unsigned f = inc >> 2, corner = inc & 3u;
            const int gs = a.fset[f];
            if (gs < 0) continue;
            const double *ge = a.fgeo + 4 * (size_t)f;
            const double nc = __ldg(ge + c), area = __ldg(ge + 3);
            const double c0 = phase_intensity(a.flab[2 * (size_t)f], a.nb_phase, a.mean), c1 = phase_intensity(a.flab[2 * (size_t)f + 1], a.nb_phase, a.mean);
            const double *fi = a.fint + 12 * (size_t)f;
            const int q0 = c_gauss_off[gs], ng = c_gauss_off[gs + 1] - q0;
            for (int k = 0; k < ng; k++) {
                const double w = c_gauss[4 * (q0 + k)], bc = c_gauss[4 * (q0 + k) + 1 + corner];
                const double fvc = nc * ((2 * __ldg(fi + k) - c0 - c1) * (c0 - c1) * w * area); // component c of Norm * magnitude (:1463-1465)
                acc += fvc * bc;
            }
        }
        a.forces[3 * (size_t)v + c] = acc;
    }
}

__global__ void __launch_bounds__(128) force_node2_kernel(const ForceNode2Args a)
{
    __shared__ double slab[16][8][12][3]; // [group in CTA][entry lane][gauss][component]
    __shared__ signed char ngs[16][8];
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const unsigned n_active = __ldg(a.n_active);
    for (unsigned item = blockIdx.x * 16 + grp; item < n_active; item += gridDim.x * 16) {
        const unsigned v = __ldg(a.active_nodes + item);
        const bool live = v >= a.begin && v < a.end;
        unsigned e0 = 0, e1 = 0;
        if (live) { e0 = __ldg(a.mesh.node_off + v); e1 = __ldg(a.mesh.node_off + v + 1); }
        double acc = 0; // lanes 0..2 of the group: x, y, z
        for (unsigned eb = e0; eb < e1; eb += 8) {
            const unsigned e = eb + lane8;
            int ng = 0;
            if (e < e1) {
                const unsigned inc = __ldg(a.mesh.node_inc + e);
                const unsigned f = inc >> 2, corner = inc & 3u;
                const int gs = a.fset[f];
                if (gs >= 0) {
                    const double *ge = a.fgeo + 4 * (size_t)f;
                    const V3 norm = {ge[0], ge[1], ge[2]};
                    const double area = ge[3];
                    const double c0 = phase_intensity(a.flab[2 * (size_t)f], a.nb_phase, a.mean), c1 = phase_intensity(a.flab[2 * (size_t)f + 1], a.nb_phase, a.mean);
                    const double *fi = a.fint + 12 * (size_t)f;
                    const int q0 = c_gauss_off[gs];
                    ng = c_gauss_off[gs + 1] - q0;
                    for (int k = 0; k < ng; k++) {
                        const double w = c_gauss[4 * (q0 + k)], bc = c_gauss[4 * (q0 + k) + 1 + corner];
                        const V3 fv = mul(norm, (2 * fi[k] - c0 - c1) * (c0 - c1) * w * area); // DEMO/segment_function.cpp:1463-1465
                        slab[grp][lane8][k][0] = fv.x * bc;
                        slab[grp][lane8][k][1] = fv.y * bc;
                        slab[grp][lane8][k][2] = fv.z * bc;
                    }
                }
            }
            ngs[grp][lane8] = (signed char)ng;
            __syncwarp(gmask);
            if (lane8 < 3) {
                const int cnt = (int)min(8u, e1 - eb);
                for (int l = 0; l < cnt; l++) {
                    const int m = ngs[grp][l];
                    for (int k = 0; k < m; k++) acc += slab[grp][l][k][lane8];
                }
            }
            __syncwarp(gmask);
        }
        if (live && lane8 < 3) a.forces[3 * (size_t)v + lane8] = acc;
    }
}

struct ForceNodeArgs {
    MeshView mesh;
    const double *fbuf;
    const signed char *fset;
    double *forces;
    unsigned begin, end;          // node range handled by this context
    const unsigned *active_nodes; // nodes that have interface faces, ascending (built with the CSR)
    const unsigned *n_active;     // device-resident length of that list
};

// 8 lanes per node: lane l fetches the addends f*coord of incident entry l (up to 12 Gauss points) into shared memory,
// then lanes 0..2 of the group add the x/y/z components in order (entries ascending = faces ascending, Gauss points in
// table order).  The loads of a node run in parallel; only the additions are sequential, as the reference's are.
__global__ void __launch_bounds__(128) force_node_kernel(const ForceNodeArgs a)
{
    __shared__ double slab[16][8][12][3]; // [group in CTA][entry lane][gauss][component]
    __shared__ signed char ngs[16][8];
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const unsigned n_active = __ldg(a.n_active);
    // grid-stride over the active nodes: the host only knows an upper bound of their number (3 per face), and a grid sized
    // for that bound is five blocks in six that exit at once
    for (unsigned item = blockIdx.x * 16 + grp; item < n_active; item += gridDim.x * 16) {
    const unsigned v = __ldg(a.active_nodes + item);
    const bool live = v >= a.begin && v < a.end;
    unsigned e0 = 0, e1 = 0;
    if (live) { e0 = __ldg(a.mesh.node_off + v); e1 = __ldg(a.mesh.node_off + v + 1); }
    double acc = 0; // lanes 0..2 of the group: x, y, z
    for (unsigned eb = e0; eb < e1; eb += 8) {
        const unsigned e = eb + lane8;
        int ng = 0;
        if (e < e1) {
            const unsigned inc = __ldg(a.mesh.node_inc + e);
            const unsigned f = inc >> 2, corner = inc & 3u;
            const int gs = a.fset[f];
            if (gs >= 0) {
                const double *fb = a.fbuf + (size_t)f * 36;
                const int q0 = c_gauss_off[gs];
                ng = c_gauss_off[gs + 1] - q0;
                for (int k = 0; k < ng; k++) {
                    const double bc = c_gauss[4 * (q0 + k) + 1 + corner];
                    slab[grp][lane8][k][0] = fb[3 * k] * bc;
                    slab[grp][lane8][k][1] = fb[3 * k + 1] * bc;
                    slab[grp][lane8][k][2] = fb[3 * k + 2] * bc;
                }
            }
        }
        ngs[grp][lane8] = (signed char)ng;
        __syncwarp(gmask);
        if (lane8 < 3) {
            const int cnt = (int)min(8u, e1 - eb);
            for (int l = 0; l < cnt; l++) {
                const int m = ngs[grp][l];
                for (int k = 0; k < m; k++) acc += slab[grp][l][k][lane8];
            }
        }
        __syncwarp(gmask);
    }
    if (live && lane8 < 3) a.forces[3 * (size_t)v + lane8] = acc;
    }
}

// Instrumentation only (bench roofline bytes): number of Gauss points the force pass evaluates, recomputed on demand so
// that the force kernels carry no same-address atomics (93 K of them serialised in L2 and cost 0.09 ms).
__global__ void gauss_count_kernel(MeshView m, unsigned begin, unsigned end, unsigned long long *count)
{
    const unsigned f = begin + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ng = 0;
    if (f < end) {
        const unsigned t0 = m.face_tets[2 * (size_t)f], t1 = m.face_tets[2 * (size_t)f + 1];
        if (t0 != kNoTet && t1 != kNoTet) {
            const double ar = tri_area(load_pos(m.pos4, m.face_nodes[3 * (size_t)f]), load_pos(m.pos4, m.face_nodes[3 * (size_t)f + 1]),
                                       load_pos(m.pos4, m.face_nodes[3 * (size_t)f + 2]));
            const int gs = gauss_set(ar);
            ng = (unsigned long long)(c_gauss_off[gs + 1] - c_gauss_off[gs]);
        }
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) ng += __shfl_xor_sync(0xffffffffu, ng, k);
    if ((threadIdx.x & 31) == 0 && ng) atomicAdd(count, ng);
}

// ---- node -> incident interface faces CSR built on the device (was a host counting sort: ~1 ms of the end-to-end step)
__global__ void csr_count_kernel(const unsigned *__restrict__ face_nodes, unsigned n_inc, unsigned n_nodes, unsigned *count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    const unsigned v = face_nodes[i];
    if (v < n_nodes) atomicAdd(count + v, 1u);
}

__global__ void csr_fill_kernel(const unsigned *__restrict__ face_nodes, unsigned n_inc, unsigned n_nodes, const unsigned *__restrict__ off, unsigned *cursor,
                                unsigned *inc)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    const unsigned v = face_nodes[i];
    if (v >= n_nodes) return;
    const unsigned slot = atomicAdd(cursor + v, 1u);
    inc[off[v] + slot] = ((i / 3u) << 2) | (i % 3u);
}

__global__ void csr_flag_kernel(const unsigned *__restrict__ count, unsigned n_nodes, unsigned *flag)
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v <= n_nodes) flag[v] = (v < n_nodes && count[v] > 0) ? 1u : 0u;
}

__global__ void csr_compact_kernel(const unsigned *__restrict__ count, const unsigned *__restrict__ apos, unsigned n_nodes, unsigned *active)
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_nodes && count[v] > 0) active[apos[v]] = v;
}

// every node's list sorted ascending (= face ascending: the reference's accumulation order)
__global__ void csr_sort_kernel(const unsigned *__restrict__ off, unsigned *inc, unsigned n_nodes)
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const unsigned b = off[v], e = off[v + 1];
    for (unsigned i = b + 1; i < e; i++) {
        const unsigned x = inc[i];
        unsigned j = i;
        while (j > b && inc[j - 1] > x) { inc[j] = inc[j - 1]; j--; }
        inc[j] = x;
    }
}

// Scatter form: one thread per interface face, per-corner sums over the Gauss points, then nine
// FP64 red.global.add per face.  Order of additions across faces is not fixed (1e-6 parity only).
template <typename VoxT> __global__ void __launch_bounds__(128) force_atomic_kernel(const ForceArgs a)
{
    const unsigned f = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    FaceGeom g;
    unsigned n[3];
    if (!face_setup(a.mesh, f, a.nb_phase, a.mean, g, n)) return;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    unsigned long long ng = 0;
    for (int q = c_gauss_off[g.gs]; q < c_gauss_off[g.gs + 1]; q++) {
        const double w = c_gauss[4 * q];
        const double b[3] = {c_gauss[4 * q + 1], c_gauss[4 * q + 2], c_gauss[4 * q + 3]};
        const V3 p = add(add(mul(g.p[0], b[0]), mul(g.p[1], b[1])), mul(g.p[2], b[2]));
        const double I = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
        const V3 fv = mul(g.norm, (2 * I - g.c0 - g.c1) * (g.c0 - g.c1) * w * g.area);
#pragma unroll
        for (int k = 0; k < 3; k++) { acc[k][0] += fv.x * b[k]; acc[k][1] += fv.y * b[k]; acc[k][2] += fv.z * b[k]; }
        ng++;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        atomicAdd(a.forces + 3 * (size_t)n[k], acc[k][0]);
        atomicAdd(a.forces + 3 * (size_t)n[k] + 1, acc[k][1]);
        atomicAdd(a.forces + 3 * (size_t)n[k] + 2, acc[k][2]);
    }
}

// =================================================================================================
// z prefix-sum table: image3d::build_sum_table (DEMO/image3d.cpp:166-176).  One thread per (x,y)
// column runs the same sequential recurrence, so the table is bit-exact; loads/stores are coalesced
// along x.
// =================================================================================================
template <typename VoxT, bool SQUARES = false> __global__ void sum_table_kernel(const VoxT *__restrict__ vol, double *__restrict__ table, int X, int Y, int Z)
{
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t XY = (size_t)X * Y;
    if (col >= XY) return;
    // SQUARES: the same table over I^2, for the data term of compute_energy (zray_resolve_kernel, mode 1); not a reference structure
    double v = decode<VoxT>(vol[col]);
    double s = SQUARES ? v * v : v;
    table[col] = s;
    for (int z = 1; z < Z; z++) {
        v = decode<VoxT>(vol[col + (size_t)z * XY]);
        s = s + (SQUARES ? v * v : v);
        table[col + (size_t)z * XY] = s;
    }
}

// =================================================================================================
// z-ray pass: update_average_intensity / compute_energy (DEMO/segment_function.cpp:1600-1764, 2331-2492)
//   1. raster (count) : per face, bounding-box pixels, 2-D barycentric inside test -> hits per (phase, column)
//   2. exclusive scan of the hit counts
//   3. raster (fill)  : same walk, hits written to their ray's segment (slot by integer atomic)
//   4. resolve        : one thread per ray restores face order, sorts with the libstdc++ algorithm,
//                       erases adjacent duplicates, pairs the hits and accumulates
// =================================================================================================
struct RayArgs {
    VolumeView vol;
    MeshView mesh;
    int nb_phase;
    unsigned face_begin, face_end;
    unsigned *ray_count;   // [nb_phase*X*Y]
    unsigned *ray_offset;  // exclusive scan of ray_count
    unsigned *ray_cursor;  // zeroed before fill
    uint2 *hits;           // (face index, packed hit)
    unsigned hits_cap;     // entries the hit buffer holds (FILL): hits beyond it are dropped and bit 1 of *overflow is raised
    int *overflow;
};

template <bool FILL> __global__ void __launch_bounds__(128) zray_raster_kernel(const RayArgs a)
{
    const unsigned f = a.face_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.face_end) return;
    const MeshView &m = a.mesh;
    if (__ldg(m.face_is_boundary + f)) return; // is_interface() and !is_boundary(), :1628
    const unsigned t0 = __ldg(m.face_tets + 2 * (size_t)f), t1 = __ldg(m.face_tets + 2 * (size_t)f + 1);
    if (t0 == kNoTet || t1 == kNoTet) return;
    const int phase0 = __ldg(m.tet_label + t0) - 1, phase1 = __ldg(m.tet_label + t1) - 1;
    const unsigned n0 = __ldg(m.face_nodes + 3 * (size_t)f), n1 = __ldg(m.face_nodes + 3 * (size_t)f + 1), n2 = __ldg(m.face_nodes + 3 * (size_t)f + 2);
    const V3 q0 = load_pos(m.pos4, n0), q1 = load_pos(m.pos4, n1), q2 = load_pos(m.pos4, n2);
    const V3 nrm = face_normal_wrt(m, n0, n1, n2, __ldg(m.tet_nodes + t0)); // get_normal(fid, tet[0]), :1639
    const bool in_1 = dot(nrm, V3{0, 0, 1}) > 0;
    const int X = a.vol.X, Y = a.vol.Y;
    const size_t XY = (size_t)X * Y;
    // bounding_box (:1578-1593) and the loop bounds floor(ld) .. round(ru) (:1644-1646), clipped to the image
    const double ldx = fmin(fmin(q0.x, q1.x), q2.x), ldy = fmin(fmin(q0.y, q1.y), q2.y);
    const double rux = fmax(fmax(q0.x, q1.x), q2.x), ruy = fmax(fmax(q0.y, q1.y), q2.y);
    const double fx0 = floor(ldx), fy0 = floor(ldy), rx1 = round(rux), ry1 = round(ruy);
    const int x0 = fx0 < 0 ? 0 : (fx0 > (double)X ? X : (int)fx0), y0 = fy0 < 0 ? 0 : (fy0 > (double)Y ? Y : (int)fy0);
    const int x1 = rx1 < 0 ? 0 : (rx1 > (double)X ? X : (int)rx1), y1 = ry1 < 0 ? 0 : (ry1 > (double)Y ? Y : (int)ry1);
    // 2-D barycentric_coords (is_mesh/util.h:283-315) with z of all operands 0
    const V3 A = {q0.x, q0.y, 0}, B = {q1.x, q1.y, 0}, C = {q2.x, q2.y, 0};
    const V3 v0 = sub(B, A), v1 = sub(C, A);
    const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1);
    const double denom = d00 * d11 - d01 * d01;
    if (fabs(denom) < 1e-8) return; // bError
    for (int x = x0; x < x1; x++)
        for (int y = y0; y < y1; y++) {
            const V3 p = {x + 0.5, y + 0.5, 0};
            const V3 v2 = sub(p, A);
            const double d20 = dot(v2, v0), d21 = dot(v2, v1);
            const double bc1 = (d11 * d20 - d01 * d21) / denom;
            const double bc2 = (d00 * d21 - d01 * d20) / denom;
            const double bc0 = 1. - bc1 - bc2;
            if (bc0 > -1e-8 && bc1 > -1e-8 && bc2 > -1e-8) {
                const double pz = (q0.z * bc0 + q1.z * bc1) + q2.z * bc2;
                const int zz = (int)nearbyint(pz);
                const size_t col = (size_t)y * X + x;
                if (phase0 >= 0 && phase0 < a.nb_phase) {
                    const size_t r = (size_t)phase0 * XY + col;
                    if (FILL) {
                        const unsigned slot = a.ray_offset[r] + atomicAdd(a.ray_cursor + r, 1u);
                        if (slot < a.hits_cap) a.hits[slot] = make_uint2(f, (unsigned)((zz << 1) | (in_1 ? 0 : 1)));
                        else atomicOr(a.overflow, 2);
                    } else atomicAdd(a.ray_count + r, 1u);
                }
                if (phase1 >= 0 && phase1 < a.nb_phase) {
                    const size_t r = (size_t)phase1 * XY + col;
                    if (FILL) {
                        const unsigned slot = a.ray_offset[r] + atomicAdd(a.ray_cursor + r, 1u);
                        if (slot < a.hits_cap) a.hits[slot] = make_uint2(f, (unsigned)((zz << 1) | (in_1 ? 1 : 0)));
                        else atomicOr(a.overflow, 2);
                    } else atomicAdd(a.ray_count + r, 1u);
                }
            }
        }
}

// ---- exclusive scan over unsigned counts (three small kernels; n up to 1024*4096 elements per level)
constexpr int kScanItems = 4096; // elements per CTA (1024 threads x 4)

__global__ void __launch_bounds__(1024) scan_block_kernel(const unsigned *__restrict__ in, unsigned *__restrict__ out, unsigned *block_sums, size_t n)
{
    __shared__ unsigned warp_sums[32];
    const size_t base = (size_t)blockIdx.x * kScanItems + (size_t)threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0u;
    const unsigned mine = v[0] + v[1] + v[2] + v[3];
    unsigned incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sums[lane];
        unsigned wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_sums[lane] = wi - w; // exclusive
        if (lane == 31 && block_sums) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    unsigned run = warp_sums[warp] + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void scan_sums_kernel(unsigned *block_sums, unsigned n_blocks, unsigned *total)
{   // single thread block, sequential carry over chunks of 1024
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned base = 0; base < n_blocks; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned mine = i < n_blocks ? block_sums[i] : 0u;
        unsigned incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warp_sums[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const unsigned excl = carry + warp_sums[warp] + incl - mine;
        if (i < n_blocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + mine;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(1024) scan_add_kernel(unsigned *out, const unsigned *block_sums, size_t n)
{
    const size_t base = (size_t)blockIdx.x * kScanItems + (size_t)threadIdx.x * 4;
    const unsigned add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (base + k < n) out[base + k] += add;
}

struct ResolveArgs {
    VolumeView vol;
    int nb_phase;
    int mode;                  // 0 = means (sum table), 1 = energy data term
    double c[kMaxPhase];       // mode 1: phase means
    const unsigned *ray_count;
    const unsigned *ray_offset;
    const double *sq_table;      // mode 1: prefix sums of I^2 along z (null: add the voxels of every interval one by one)
    uint2 *hits;                 // rays with more than kMaxRayHits hits are sorted in place
    unsigned hits_cap;
    unsigned col_begin, col_end; // column range of this partition
    double *partial_sum;         // [nb_phase][gridDim.x]
    long long *partial_cnt;      // [nb_phase][gridDim.x]
    int *overflow;
};

template <typename VoxT> __global__ void __launch_bounds__(128) zray_resolve_kernel(const ResolveArgs a)
{
    const int phase = blockIdx.y;
    const int X = a.vol.X, Y = a.vol.Y, Z = a.vol.Z;
    const size_t XY = (size_t)X * Y;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    double sum = 0;
    long long cnt = 0;
    const unsigned col = a.col_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (col < a.col_end) {
        const size_t r = (size_t)phase * XY + col;
        const unsigned n_hits = __ldg(a.ray_count + r);
        const unsigned off = __ldg(a.ray_offset + r);
        if (n_hits > 1 && off + n_hits <= a.hits_cap) { // rays with a single hit are ignored (:1701); a ray past the buffer was not recorded (overflow bit 1 is up)
            unsigned fidx[kMaxRayHits];
            int hloc[kMaxRayHits];
            uint2 *src = a.hits + off;
            int n = (int)n_hits;
            int *h = hloc;
            if (n_hits <= (unsigned)kMaxRayHits) {
                // restore the reference's push order = ascending face key (the fill order is arbitrary)
                for (int i = 0; i < n; i++) {
                    const uint2 e = src[i];
                    int j = i - 1;
                    while (j >= 0 && fidx[j] > e.x) { fidx[j + 1] = fidx[j]; hloc[j + 1] = hloc[j]; j--; }
                    fidx[j + 1] = e.x; hloc[j + 1] = (int)e.y;
                }
            } else {
                // a ray through more interface faces than the local arrays hold (the reference has no limit): the same steps in place,
                // on the ray's own stretch of the hit buffer -- slow, and rare
                for (int i = 1; i < n; i++) {
                    const uint2 e = src[i];
                    int j = i - 1;
                    while (j >= 0 && src[j].x > e.x) { src[j + 1] = src[j]; j--; }
                    src[j + 1] = e;
                }
                h = reinterpret_cast<int *>(src);
                for (int i = 0; i < n; i++) h[i] = (int)src[i].y; // entry i is read before words 2i, 2i+1 >= i are overwritten
            }
            std_sort_hits(h, n);
            n = erase_adjacent_duplicates(h, n);
            if ((n & 1) == 0) { // odd rays are silently dropped (:1721, 1741)
                const int x = (int)(col % X), y = (int)(col / X);
                for (int j = 0; j < n / 2; j++) {
                    int z1 = h[2 * j] >> 1, z2 = h[2 * j + 1] >> 1;
                    if (z1 < 0) z1 = 0;
                    if (z2 < z1) z2 = z1;
                    cnt += z2 - z1;
                    if (a.mode == 0) {
                        int lo = z1, hi = z2; // image3d::sum_line_z clamps (image3d.cpp:193-201)
                        if (hi >= Z) { hi = Z - 1; if (lo >= Z) lo = Z - 1; }
                        sum += a.vol.sum_table[(size_t)hi * XY + col] - a.vol.sum_table[(size_t)lo * XY + col];
                    } else if (a.sq_table) {
                        // sum over k in (z1, z2) of (I_k - c)^2 = S2 - 2 c S1 + c^2 n from the two prefix tables (1e-6 quantity: the
                        // reference adds the squares one by one, :2470-2473); voxels past the last slice read 0 (image3d.cpp:480-491)
                        const double cc = a.c[phase];
                        const int n_all = z2 - 1 - z1; // k = z1 + 1 .. z2 - 1
                        if (n_all > 0) {
                            const int hi = min(z2 - 1, Z - 1);
                            double s1 = 0, s2 = 0;
                            int n_in = 0;
                            if (hi > z1) {
                                n_in = hi - z1;
                                s1 = a.vol.sum_table[(size_t)hi * XY + col] - a.vol.sum_table[(size_t)z1 * XY + col];
                                s2 = a.sq_table[(size_t)hi * XY + col] - a.sq_table[(size_t)z1 * XY + col];
                            }
                            sum += (s2 - 2.0 * cc * s1) + cc * cc * (double)n_all;
                            (void)n_in;
                        }
                    } else {
                        const double cc = a.c[phase];
                        for (int k = z1 + 1; k < z2; k++) {
                            const double d = fetch(vol, X, Y, Z, x, y, k) - cc;
                            sum += d * d; // std::pow(., 2)
                        }
                    }
                }
            }
        }
    }
    // CTA reduction -> one partial per (phase, CTA); finalize adds them in CTA order
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        sum += shfl_xor_d(sum, m);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
    }
    __shared__ double ssum[4];
    __shared__ long long scnt[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        long long c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { s += ssum[w]; c += scnt[w]; }
        a.partial_sum[(size_t)phase * gridDim.x + blockIdx.x] = s;
        a.partial_cnt[(size_t)phase * gridDim.x + blockIdx.x] = c;
    }
}

// sums[0..P) = total, [P..2P) = 0 (no squared sum on this path), [2P..3P) = voxel count.  One warp per phase: lane l adds
// partials l, l+32, ... in order, then a fixed xor tree (deterministic; counts are integers and exact in any order).
__global__ void zray_finalize_kernel(const double *partial_sum, const long long *partial_cnt, int n_blocks, int nb_phase, double *sums, double *mean)
{
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= nb_phase) return;
    double s = 0;
    long long c = 0;
    for (int i = lane; i < n_blocks; i += 32) { s += partial_sum[(size_t)k * n_blocks + i]; c += partial_cnt[(size_t)k * n_blocks + i]; }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        s += shfl_xor_d(s, m);
        c += __shfl_xor_sync(0xffffffffu, c, m);
    }
    if (lane == 0) {
        sums[k] = s;
        sums[nb_phase + k] = 0;
        sums[2 * nb_phase + k] = (double)c;
        if (mean) mean[k] = s / (double)c;
    }
}

// Snapshot index validation on the device: out-of-range node keys / tet indices are neutralised (0 / NO_TET) so that no
// later kernel can read out of bounds, and *flag is raised; the host reports the error at its next synchronisation.
__global__ void validate_snapshot_kernel(unsigned *tet_nodes, size_t n_tn, unsigned *face_nodes, size_t n_fn, unsigned *face_tets, size_t n_ft,
                                         unsigned n_nodes, unsigned n_tets, int *flag)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tn + n_fn + n_ft; i += stride) {
        if (i < n_tn) {
            if (tet_nodes[i] >= n_nodes) { tet_nodes[i] = 0; bad = true; }
        } else if (i < n_tn + n_fn) {
            if (face_nodes[i - n_tn] >= n_nodes) { face_nodes[i - n_tn] = 0; bad = true; }
        } else {
            const unsigned t = face_tets[i - n_tn - n_fn];
            if (t >= n_tets && t != kNoTet) { face_tets[i - n_tn - n_fn] = kNoTet; bad = true; }
        }
    }
    if (bad) atomicExch(flag, 1);
}

// Incremental snapshot: moved nodes and changed tets scattered into the resident arrays.  Entries with a
// key / index out of range are dropped and *flag is raised, like validate_snapshot_kernel.
__global__ void scatter_positions_kernel(const unsigned *__restrict__ keys, const double *__restrict__ pos3, unsigned n, double4 *pos4, float4 *pos4f,
                                         unsigned n_nodes, int *flag)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys ? keys[i] : i;
    if (k >= n_nodes) { atomicExch(flag, 1); return; }
    const double x = pos3[3 * (size_t)i], y = pos3[3 * (size_t)i + 1], z = pos3[3 * (size_t)i + 2];
    pos4[k] = make_double4(x, y, z, 0.0);
    pos4f[k] = make_float4((float)x, (float)y, (float)z, 0.f);
}

__global__ void scatter_tets_kernel(const unsigned *__restrict__ idx, const uint4 *__restrict__ nodes, const int *__restrict__ labels, unsigned n, uint4 *tet_nodes,
                                    int *tet_label, unsigned n_tets, unsigned n_nodes, int *flag)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned t = idx[i];
    const uint4 nd = nodes[i];
    if (t >= n_tets || nd.x >= n_nodes || nd.y >= n_nodes || nd.z >= n_nodes || nd.w >= n_nodes) { atomicExch(flag, 1); return; }
    tet_nodes[t] = nd;
    tet_label[t] = labels[i];
}

// m_vertex_bound of update_vertex_boundary (DEMO/segment_function.cpp:2158-2171) from the snapshot
__global__ void vertex_bound_kernel(MeshView m, unsigned char *vb)
{
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.n_faces) return;
    const unsigned t0 = m.face_tets[2 * (size_t)f], t1 = m.face_tets[2 * (size_t)f + 1];
    if (t0 == kNoTet || t1 == kNoTet) return;
    if (m.tet_label[t0] == 0 || m.tet_label[t1] == 0) {
        vb[m.face_nodes[3 * (size_t)f]] = 1;
        vb[m.face_nodes[3 * (size_t)f + 1]] = 1;
        vb[m.face_nodes[3 * (size_t)f + 2]] = 1;
    }
}

// =================================================================================================
// Initialisation pass of segment_function::initialization_discrete_opt (DEMO/segment_function.cpp:165-262), SURVEY 8(f) row 2:
// histogram of the per-tet mean intensities (bin (int)(mean*256), clamped to 255, :219-226) for Otsu's multi-level
// thresholding (host: DEMO/otsu_multi.h), then label = 1 + lower_bound(thresholds, min((int)(mean*255), 255)) (:250-260).
// Label-0 (boundary layer) tets are skipped by both (:167-168, :245-248).
// =================================================================================================
// Per-tet mean of the initialisation pass with the reference's own order of additions: get_tetra_intensity adds the sampled voxels one
// after another in table order (DEMO/image3d.cpp:363-367), then total * v / n, then / v (:371, :381).  The histogram bin and the label
// are truncations of that mean, so the sum must round like the reference's -- voxels of integer volumes (q / 255.0, q / 65535.0) do
// not add exactly.  One thread per tet, sequential: the pass runs once per segmentation.
template <typename VoxT>
__global__ void __launch_bounds__(128) tet_mean_ordered_kernel(VolumeView vv, MeshView m, const double *__restrict__ tet_bary, double *__restrict__ tet_mean)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m.n_tets) return;
    if (__ldg(m.tet_label + t) <= 0) { tet_mean[t] = -1.0; return; }
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(vv.data);
    const uint4 nd = __ldg(m.tet_nodes + t);
    const V3 p0 = load_pos(m.pos4, nd.x), p1 = load_pos(m.pos4, nd.y), p2 = load_pos(m.pos4, nd.z), p3 = load_pos(m.pos4, nd.w);
    const double v = tet_volume(p0, p1, p2, p3);
    const int L = tet_level(v);
    const int rows = (L + 1) * (L + 1) * (L + 1);
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(tet_bary) + 2 * (size_t)((L * (L + 1) / 2) * (L * (L + 1) / 2));
    double total = 0;
    for (int i = 0; i < rows; i++) {
        const double2 b01 = __ldg(tb + 2 * i), b23 = __ldg(tb + 2 * i + 1);
        const int ix = (int)(((p0.x * b01.x + p1.x * b01.y) + p2.x * b23.x) + p3.x * b23.y);
        const int iy = (int)(((p0.y * b01.x + p1.y * b01.y) + p2.y * b23.x) + p3.y * b23.y);
        const int iz = (int)(((p0.z * b01.x + p1.z * b01.y) + p2.z * b23.x) + p3.z * b23.y);
        total += fetch(vol, vv.X, vv.Y, vv.Z, ix, iy, iz);
    }
    total = total * v / (double)rows;
    tet_mean[t] = total / v;
}

__global__ void __launch_bounds__(256) init_hist_kernel(const int *__restrict__ tet_label, const double *__restrict__ tet_mean, unsigned n_tets, int *hist)
{
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tets; t += gridDim.x * blockDim.x) {
        if (tet_label[t] <= 0) continue;
        int idx = (int)(tet_mean[t] * 256);
        if (idx > 255) idx = 255;
        if (idx < 0) idx = 0; // a negative mean (negative voxels) would index out of bounds in the reference
        atomicAdd(&sh[idx], 1);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void init_label_kernel(const int *__restrict__ tet_label, const double *__restrict__ tet_mean, unsigned n_tets, int n_thres, const int *__restrict__ thres,
                                  int *__restrict__ out)
{
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tets) return;
    if (tet_label[t] <= 0) { out[t] = tet_label[t] < 0 ? tet_label[t] : 0; return; }
    int m = (int)(tet_mean[t] * 255);
    if (m >= 255) m = 255;
    int k = 0;
    while (k < n_thres && thres[k] < m) k++; // std::lower_bound on the ascending thresholds
    out[t] = k + 1;
}

// =================================================================================================
// Internal (curvature) force: segment_function::compute_internal_force_2 (DEMO/segment_function.cpp:2197-2232), SURVEY 8(f) row 3.
// Two kernels like the external force, so that every node adds its terms in the reference's order (faces ascending) without
// atomics:
//   iforce_face_kernel  thread per interface face: the three corner terms -h |l2 - l1|, h = normalize(l0 - proj(l0 on l1 l2))
//                       (is_mesh/util.h:395-399), skipped when all three nodes are vertex-bound (segment_function::is_boundary(
//                       FaceKey), :2186-2196), and the face normal get_normal(fid) (src/DSC.h:3129-3133) for the node normals;
//   iforce_node_kernel  thread per node that has interface faces: ordered sums over its CSR entries, then for interface,
//                       non-crossing nodes with |f| > EPSILON the projection f = n dot(n, f) on the node normal
//                       (DSC::get_normal(node), src/DSC.h:3138-3159).  The reference adds the face normals in get_faces(node)
//                       order (is_mesh/is_mesh.h:691-699: edges of the node, faces of each edge, first occurrence): topology the
//                       flat snapshot does not carry.  With ord_off / ord_idx (dscgpu_set_node_face_order) the normal is summed in
//                       that order and every force is bit-identical; without, in ascending face order: the unprojected sums are
//                       bit-identical, the projected ones agree to rounding.
// =================================================================================================
__global__ void __launch_bounds__(128) iforce_face_kernel(MeshView m, const unsigned char *__restrict__ vb, double *__restrict__ fterm, double *__restrict__ fnorm,
                                                          unsigned char *__restrict__ fskip)
{
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.n_faces) return;
    const unsigned n0 = __ldg(m.face_nodes + 3 * (size_t)f), n1 = __ldg(m.face_nodes + 3 * (size_t)f + 1), n2 = __ldg(m.face_nodes + 3 * (size_t)f + 2);
    const unsigned t0 = __ldg(m.face_tets + 2 * (size_t)f), t1 = __ldg(m.face_tets + 2 * (size_t)f + 1);
    // get_normal(fid): sorted w.r.t. the co-boundary tet with the highest label, first wins (is_mesh/is_mesh.h:639-663)
    unsigned tid = t0;
    if (t0 == kNoTet) tid = t1;
    else if (t1 != kNoTet && __ldg(m.tet_label + t1) > __ldg(m.tet_label + t0)) tid = t1;
    V3 N = {0, 0, 0};
    if (tid != kNoTet) N = face_normal_wrt(m, n0, n1, n2, __ldg(m.tet_nodes + tid));
    fnorm[3 * (size_t)f] = N.x; fnorm[3 * (size_t)f + 1] = N.y; fnorm[3 * (size_t)f + 2] = N.z;
    const bool skip = vb[n0] && vb[n1] && vb[n2];
    fskip[f] = skip ? 1 : 0;
    if (skip) return;
    const V3 p[3] = {load_pos(m.pos4, n0), load_pos(m.pos4, n1), load_pos(m.pos4, n2)};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const V3 l0 = p[i], l1 = p[(i + 1) % 3], l2 = p[(i + 2) % 3];
        const V3 v = sub(l2, l1), v1 = sub(l0, l1);
        const V3 proj = add(l1, divs(mul(v, dot(v1, v)), dot(v, v))); // a + v * dot(v1,v) / dot(v,v)
        const V3 h = normalize(sub(l0, proj));
        const V3 t = mul(mul(h, -1.0), length(v)); // -h * (l2 - l1).length()
        double *o = fterm + 9 * (size_t)f + 3 * i;
        o[0] = t.x; o[1] = t.y; o[2] = t.z;
    }
}

__global__ void __launch_bounds__(128) iforce_node_kernel(MeshView m, const unsigned *__restrict__ active_nodes, const unsigned *__restrict__ n_active,
                                                          const unsigned char *__restrict__ node_flags, const double *__restrict__ fterm,
                                                          const double *__restrict__ fnorm, const unsigned char *__restrict__ fskip,
                                                          const unsigned *__restrict__ ord_off, const unsigned *__restrict__ ord_idx, double *__restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_active) return;
    const unsigned n = active_nodes[i];
    V3 F = {0, 0, 0}, S = {0, 0, 0};
    for (unsigned k = m.node_off[n]; k < m.node_off[n + 1]; k++) {
        const unsigned e = m.node_inc[k], f = e >> 2, c = e & 3u;
        S = add(S, V3{fnorm[3 * (size_t)f], fnorm[3 * (size_t)f + 1], fnorm[3 * (size_t)f + 2]});
        if (!fskip[f]) {
            const double *t = fterm + 9 * (size_t)f + 3 * c;
            F = add(F, V3{t[0], t[1], t[2]});
        }
    }
    const unsigned char fl = node_flags[n];
    if ((fl & 1) && !(fl & 2) && length(F) > 1e-8) { // is_interface && !is_crossing && |f| > EPSILON (is_mesh/util.h:43)
        if (ord_off) { // the caller's get_faces(node) order
            S = V3{0, 0, 0};
            for (unsigned k = ord_off[n]; k < ord_off[n + 1]; k++) {
                const size_t f = ord_idx[k];
                S = add(S, V3{fnorm[3 * f], fnorm[3 * f + 1], fnorm[3 * f + 2]});
            }
        }
        V3 norm = {0, 0, 0};
        if (!(length(S) < 1e-8)) norm = normalize(S);
        F = mul(norm, dot(norm, F));
    }
    out[3 * (size_t)n] = F.x; out[3 * (size_t)n + 1] = F.y; out[3 * (size_t)n + 2] = F.z;
}

// alpha * sum of areas of interface faces whose nodes are not all vertex-bound (:2482-2489);
// one partial per CTA, added in CTA order by the caller's finalize.
__global__ void __launch_bounds__(256) area_term_kernel(MeshView m, const unsigned char *vb, double alpha, double *partials)
{
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (f < m.n_faces) {
        const unsigned n0 = m.face_nodes[3 * (size_t)f], n1 = m.face_nodes[3 * (size_t)f + 1], n2 = m.face_nodes[3 * (size_t)f + 2];
        if (!(vb[n0] && vb[n1] && vb[n2])) s = alpha * tri_area(load_pos(m.pos4, n0), load_pos(m.pos4, n1), load_pos(m.pos4, n2));
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) s += shfl_xor_d(s, k);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sm[w];
        partials[blockIdx.x] = r;
    }
}

__global__ void sum_partials_kernel(const double *partials, int n, double *out)
{   // one warp: lane l adds partials l, l+32, ... in order, then a fixed xor tree
    const int lane = threadIdx.x & 31;
    double r = 0;
    for (int i = lane; i < n; i += 32) r += partials[i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) r += shfl_xor_d(r, m);
    if (threadIdx.x == 0) *out = r;
}

// =================================================================================================
// SURVEY 8(f) row 3, second half -- the non-topological part of segment_function::snapp_boundary (DEMO/segment_function.cpp:466-541)
// with get_node_displacement (:581-584): the external and internal forces stay on the device until the destinations are formed.
//   stage 1 (node_displacement_kernel): for every node with interface faces (is_interface() or is_crossing(), :475-476)
//           dis = (F + alpha Fi) dt, destination = pos + dis; nodes of m_vertex_bound are aligned to the faces of the image domain
//           (macro algin_pos, :59-65); node_displacement = destination - pos; the largest length over the free nodes and over the
//           bound nodes (maxima: exact in any order).
//   stage 2 (node_destination_kernel): scale = max_displacement / max_dis when max_dis exceeds it (:512-519); bound nodes are clipped to
//           max_dis (:531-534); destination = pos + dis * dt_adapt (:536); _cur_dis = dis; the largest |dis| dt_adapt (:540).
// Everything is element-wise FP64 in the reference's order: bit-identical results.
// =================================================================================================
struct NodeMoveArgs {
    const double4 *pos4;
    const double *forces, *iforces;   // 3 per node
    const unsigned char *vbound;      // m_vertex_bound
    const unsigned char *node_flags;  // bit 0 is_interface, bit 1 is_crossing (null: every node of active_nodes counts)
    const double *dt_adapt;           // per node, null = 1.0
    const unsigned *active_nodes, *n_active;
    double alpha, dt, threshold, dom[3], max_displacement;
    double *disp;                     // 3 per node: node_displacement, then (stage 2) _cur_dis
    double *dest;                     // 3 per node: destinations (nodes that do not move keep their position)
    unsigned long long *maxbits;      // [0] max_dis, [1] max_boundary_dis, [2] real_max as bit patterns of non-negative doubles
    unsigned n_nodes;
};

__global__ void __launch_bounds__(256) node_displacement_kernel(const NodeMoveArgs a)
{
    const unsigned n_active = __ldg(a.n_active);
    double m_free = 0, m_bound = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_active; i += gridDim.x * blockDim.x) {
        const unsigned v = __ldg(a.active_nodes + i);
        if (a.node_flags && !(a.node_flags[v] & 3)) continue;
        const V3 pos = load_pos(a.pos4, v);
        const double *F = a.forces + 3 * (size_t)v, *Fi = a.iforces + 3 * (size_t)v;
        const double dis[3] = {(F[0] + a.alpha * Fi[0]) * a.dt, (F[1] + a.alpha * Fi[1]) * a.dt, (F[2] + a.alpha * Fi[2]) * a.dt};
        const double p[3] = {pos.x, pos.y, pos.z};
        double d[3] = {p[0] + dis[0], p[1] + dis[1], p[2] + dis[2]};
        const bool bound = a.vbound[v] != 0;
        if (bound) {
#pragma unroll
            for (int k = 0; k < 3; k++) { // algin_pos(k)
                if (p[k] < a.threshold) d[k] = 0;
                if (p[k] > a.dom[k] - 1 - a.threshold) d[k] = a.dom[k] - 1;
                if (d[k] < a.threshold) d[k] = 0;
                if (d[k] > a.dom[k] - 1 - a.threshold) d[k] = a.dom[k] - 1;
            }
        }
        const V3 nd = {d[0] - p[0], d[1] - p[1], d[2] - p[2]};
        const double len = length(nd);
        if (bound) m_bound = fmax(m_bound, len); else m_free = fmax(m_free, len);
        double *o = a.disp + 3 * (size_t)v;
        o[0] = nd.x; o[1] = nd.y; o[2] = nd.z;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { m_free = fmax(m_free, shfl_xor_d(m_free, m)); m_bound = fmax(m_bound, shfl_xor_d(m_bound, m)); }
    if ((threadIdx.x & 31) == 0) { // non-negative doubles order like their bit patterns
        if (m_free > 0) atomicMax(a.maxbits, (unsigned long long)__double_as_longlong(m_free));
        if (m_bound > 0) atomicMax(a.maxbits + 1, (unsigned long long)__double_as_longlong(m_bound));
    }
}

__global__ void __launch_bounds__(256) node_destination_kernel(const NodeMoveArgs a)
{
    const unsigned n_active = __ldg(a.n_active);
    double max_dis = __longlong_as_double((long long)a.maxbits[0]);
    double scale = 1.0;
    if (max_dis > a.max_displacement) { scale = a.max_displacement / max_dis; max_dis = a.max_displacement; }
    double real_max = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_active; i += gridDim.x * blockDim.x) {
        const unsigned v = __ldg(a.active_nodes + i);
        if (a.node_flags && !(a.node_flags[v] & 1)) continue; // the second loop takes is_interface() nodes only (:527)
        double *o = a.disp + 3 * (size_t)v;
        V3 dis = {o[0] * scale, o[1] * scale, o[2] * scale};
        const double len0 = length(dis);
        if (a.vbound[v] && len0 > max_dis) dis = mul(dis, max_dis / len0);
        const double ad = a.dt_adapt ? a.dt_adapt[v] : 1.0;
        const V3 pos = load_pos(a.pos4, v);
        double *dst = a.dest + 3 * (size_t)v;
        dst[0] = pos.x + dis.x * ad; dst[1] = pos.y + dis.y * ad; dst[2] = pos.z + dis.z * ad;
        o[0] = dis.x; o[1] = dis.y; o[2] = dis.z;
        real_max = fmax(real_max, length(dis) * ad);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) real_max = fmax(real_max, shfl_xor_d(real_max, m));
    if ((threadIdx.x & 31) == 0 && real_max > 0) atomicMax(a.maxbits + 2, (unsigned long long)__double_as_longlong(real_max));
}

// dest = pos for every node (the nodes stage 2 does not touch keep their position, like is_mesh's destination after deform)
__global__ void node_dest_init_kernel(const double4 *__restrict__ pos4, double *__restrict__ dest, double *__restrict__ disp, unsigned n)
{
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const V3 p = load_pos(pos4, v);
    dest[3 * (size_t)v] = p.x; dest[3 * (size_t)v + 1] = p.y; dest[3 * (size_t)v + 2] = p.z;
    disp[3 * (size_t)v] = 0; disp[3 * (size_t)v + 1] = 0; disp[3 * (size_t)v + 2] = 0;
}

// =================================================================================================
// A9b -- the force loop of segment_function::thickenning_surface (DEMO/segment_function.cpp:1990-2051): the quadrature of
// compute_external_force over the interface faces that are not segment_function::is_boundary (all three nodes in m_vertex_bound,
// :2186-2196), plus, per face, the mean and the mean absolute deviation of the scalar magnitudes
// (2I - c0 - c1)(c0 - c1) w area of its Gauss points (:2030-2043) that decide which edges get split (:2046-2049, host).
// This kernel masks the skipped faces (fset_out = -1, so that force_node2_kernel leaves them out) and forms the statistics in the
// reference's order: sum in Gauss-point order, divide, sum of |m - mean|, divide.
// =================================================================================================
__global__ void __launch_bounds__(128) face_force_stats_kernel(MeshView m, const double *__restrict__ fgeo, const int *__restrict__ flab, const double *__restrict__ fint,
                                                               const signed char *__restrict__ fset, const unsigned char *__restrict__ vb, int nb_phase,
                                                               const double *__restrict__ mean, signed char *__restrict__ fset_out, double *__restrict__ face_mean,
                                                               double *__restrict__ face_dev)
{
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.n_faces) return;
    int gs = fset[f];
    if (gs >= 0) {
        const unsigned n0 = m.face_nodes[3 * (size_t)f], n1 = m.face_nodes[3 * (size_t)f + 1], n2 = m.face_nodes[3 * (size_t)f + 2];
        if (vb[n0] && vb[n1] && vb[n2]) gs = -1;
    }
    fset_out[f] = (signed char)gs;
    double mu = 0, dev = 0;
    if (gs >= 0) {
        const double area = fgeo[4 * (size_t)f + 3];
        const double c0 = phase_intensity(flab[2 * (size_t)f], nb_phase, mean), c1 = phase_intensity(flab[2 * (size_t)f + 1], nb_phase, mean);
        const int q0 = c_gauss_off[gs], ng = c_gauss_off[gs + 1] - q0;
        double mag[12];
        for (int k = 0; k < ng; k++) {
            mag[k] = (2 * fint[12 * (size_t)f + k] - c0 - c1) * (c0 - c1) * c_gauss[4 * (q0 + k)] * area;
            mu += mag[k];
        }
        mu /= (double)ng;
        for (int k = 0; k < ng; k++) dev += fabs(mag[k] - mu);
        dev /= (double)ng;
    }
    face_mean[f] = mu;
    face_dev[f] = dev;
}

// rows[i] = forces[active_nodes[i]]: the rows of the force array that can be non-zero, for the compact read-back
__global__ void gather_rows_kernel(const double *__restrict__ forces, const unsigned *__restrict__ active_nodes, unsigned n, double *__restrict__ rows)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t v = active_nodes[i];
    rows[3 * (size_t)i] = forces[3 * v]; rows[3 * (size_t)i + 1] = forces[3 * v + 1]; rows[3 * (size_t)i + 2] = forces[3 * v + 2];
}

// The whole result of an evaluation in one buffer, so that one copy brings it to the host: header [3 kMaxPhase sums | kMaxPhase means],
// then rows[i] = forces[active_nodes[i]], then the keys active_nodes[0..n).
__global__ void gather_result_kernel(const double *__restrict__ forces, const unsigned *__restrict__ active_nodes, unsigned n, const double *__restrict__ sums,
                                     const double *__restrict__ mean, int nb_phase, double *__restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (unsigned)(3 * nb_phase)) out[i] = sums[i];
    if (i < (unsigned)nb_phase) out[3 * kMaxPhase + i] = mean[i];
    if (i >= n) return;
    double *rows = out + 4 * kMaxPhase;
    unsigned *keys = reinterpret_cast<unsigned *>(rows + 3 * (size_t)n);
    const unsigned v = active_nodes[i];
    rows[3 * (size_t)i] = forces[3 * (size_t)v]; rows[3 * (size_t)i + 1] = forces[3 * (size_t)v + 1]; rows[3 * (size_t)i + 2] = forces[3 * (size_t)v + 2];
    keys[i] = v;
}

// dst = src on the stream (a 4-byte device-to-device copy would queue on a copy engine behind the snapshot upload)
__global__ void copy_u32_kernel(unsigned *dst, const unsigned *src) { *dst = *src; }

} // namespace dsc
