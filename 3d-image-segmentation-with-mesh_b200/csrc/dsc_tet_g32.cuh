// dsc_tet_g32.cuh -- per-tet integration, warp-per-tet form over a bricked copy of the volume (variant 1024, the default
// of the per-iteration hot path).
//
// Same arithmetic contract as tet_integrate_w2_kernel (dsc_kernels.cuh): image3d::get_tetra_intensity
// (DEMO/image3d.cpp:351-382) batched over all tets and fused with the per-label reduction; the sample -> voxel
// classification is bit-identical to the reference (FP32 lattice filter with a rigorous guard band, exact FP64 re-evaluation
// inside the band), sums are FP64.  What changes is everything the profile of w2 showed to be the limit
// (profiles/README.md: the L1 data pipe at one wavefront per distinct 128 B line touched by a gather -- 16.8 lines per
// 32-lane gather on the x-fastest volume -- and instruction issue at ~52 instructions per sample):
//
//  * bricked volume: the kernel gathers from a copy of the volume laid out in 4x2x4-voxel bricks (32 voxels; 128 B for
//    float), padded by 8 voxels on every side.  A tet's samples fall into few bricks, and with all 32 lanes on ONE tet a
//    gather touches 8.0 lines instead of 16.8 (scripts/line_sim.py reproduces both figures on the C4 mesh).  The padding
//    holds what image3d::get_value returns outside the volume (0, DEMO/image3d.cpp:480-491) and, in the layer at -1, a
//    replica of layer 0: the reference truncates toward zero (image3d.cpp:366), so a coordinate in (-1,0) reads voxel 0.
//    Every tet whose bounding box lies inside the padded region therefore takes the fast path without bounds checks.
//  * tet-local origins are aligned to brick corners, so the brick index and the offset inside the brick are both linear
//    in r and floor(r / brick): idx = base + rx + 28 hx + 4 ry + (Sy-8) hy + 8 rz + (Sz-32) hz.
//  * packed FP32 (Blackwell FFMA2 / FADD2, PTX fma.rn.f32x2): one lane evaluates two samples per instruction -- the two
//    halves of a pair of tets (pairs are formed in phase A from tets of the same label and sample count) -- which halves
//    the issue slots of the position arithmetic.  floor() is the 1.5*2^23 trick; floor(r/4) and floor(r/2) come from one
//    more FFMA2 each on the already rounded value: fma(t, 1-2^-23, 4.5*2^23) rounds M4 + r - 1.5 - r 2^-23 at ulp 4.
//  * no per-tet cross-lane reduction: each lane multiplies its partial sum by the tet's v/n^3 and keeps a running sum for
//    the current label; a warp reduction happens only when the label changes (tets are sorted by label inside a tile).
//
// Lattice form of the sample position and the guard band: see tet_integrate_w2_kernel; the only difference is that the
// origin O is floor(min corner) rounded DOWN to a brick corner, so Rmax is up to 3 voxels larger.
#pragma once
#include <type_traits>
#include "dsc_kernels.cuh"

#ifndef G32_MIN_BLOCKS
#define G32_MIN_BLOCKS 2
#endif

namespace dsc {

constexpr int kBrickPad = 8;                      // voxels of padding on every side (multiple of every brick extent)
constexpr int kG32Slots = 66;                     // sum over n = 1..9 of ceil(n^3 / 32)
constexpr int kG32Rows = kG32Slots * 32;          // padded table rows
constexpr int kG32WarpBytes = 2048 + 256 + 256 + 512 + 256 + 128 + 128 + 256 + 128; // rec, w, control words, nd, v, meta, base, slow w, slow meta
constexpr int kG32SmemBytes = kG32Rows * 16 + 8 * kG32WarpBytes;

// first 32-row slot of level L (n = L + 1) in the padded table
DSC_HD int g32_slot_off(int L)
{
    const int off[10] = {0, 1, 2, 3, 5, 9, 16, 27, 43, 66};
    return off[L];
}

// element index of voxel (x, y, z) (volume coordinates, -kBrickPad <= x < X + pad ...) in the bricked, padded copy
DSC_HD unsigned brick_index(int x, int y, int z, unsigned Sy, unsigned Sz)
{
    const unsigned xp = (unsigned)(x + kBrickPad), yp = (unsigned)(y + kBrickPad), zp = (unsigned)(z + kBrickPad);
    return (zp >> 2) * Sz + (yp >> 1) * Sy + ((xp >> 2) << 5) + ((zp & 3u) << 3) + ((yp & 1u) << 2) + (xp & 3u);
}

// Builds the bricked copy.  flags bit 0 is raised when a float volume holds anything but non-negative finite numbers
// (negative values incl. -0, infinities, NaNs): the kernel then widens voxels with the general routine.
template <typename VoxT>
__global__ void brick_volume_kernel(const VoxT *__restrict__ vol, VoxT *__restrict__ out, int X, int Y, int Z, int NBX, int NBY, int NBZ, unsigned *flags)
{
    const size_t n = (size_t)NBX * NBY * NBZ * 32;
    unsigned f = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned inner = (unsigned)(i & 31);
        size_t b = i >> 5;
        const int bx = (int)(b % NBX);
        b /= NBX;
        const int by = (int)(b % NBY);
        const int bz = (int)(b / NBY);
        const int x = bx * 4 + (int)(inner & 3) - kBrickPad, y = by * 2 + (int)((inner >> 2) & 1) - kBrickPad, z = bz * 4 + (int)(inner >> 3) - kBrickPad;
        VoxT v = (VoxT)0;
        if (x >= -1 && x < X && y >= -1 && y < Y && z >= -1 && z < Z) {
            v = vol[((size_t)max(z, 0) * Y + max(y, 0)) * X + max(x, 0)];
            if constexpr (sizeof(VoxT) == 4 && !VoxTraits<VoxT>::integral) {
                const unsigned bits = __float_as_uint((float)v);
                if (bits >= 0x7f800000u) f = 1; // negative (sign bit), infinite or NaN
            }
        }
        out[i] = v;
    }
    if (f) atomicOr(flags, 1u);
}

struct TetG32Args {
    MeshView mesh;
    const double *tet_bary;  // exact table rows (fallback inside the guard band, slow tets)
    const float4 *lat32;     // lattice rows (a1, a2, a3, eps), every level padded to a multiple of 32 rows
    const void *brick;       // bricked, padded copy of the volume
    unsigned Sy, Sz;         // brick strides in elements: 32 * NBX, Sy * NBY
    int X, Y, Z;
    unsigned tet_begin, tet_end;
    int nb_phase;
    double *tet_total, *tet_vol, *tet_mean;
    double *partials;
    unsigned long long *sample_count;
};

// ---- packed FP32 pairs (Blackwell FFMA2 / FADD2) ----
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

DSC_D double warp_sum_d(double v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
    return v;
}

// exact voxel of sample row i of a tet (get_coord, DEMO/tet_dis_coord.hpp:27; truncation image3d.cpp:366; bounds
// image3d.cpp:480-491) as an index into the bricked copy; samples outside the volume map to element 0, which is padding (0)
DSC_D unsigned g32_exact_index(const TetG32Args &a, V3 p0, V3 p1, V3 p2, V3 p3, int L, int i)
{
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(a.tet_bary) + 2 * ((size_t)((L * (L + 1) / 2) * (L * (L + 1) / 2)) + (size_t)i);
    const double2 b01 = __ldg(tb), b23 = __ldg(tb + 1);
    const int ix = (int)(((p0.x * b01.x + p1.x * b01.y) + p2.x * b23.x) + p3.x * b23.y);
    const int iy = (int)(((p0.y * b01.x + p1.y * b01.y) + p2.y * b23.x) + p3.y * b23.y);
    const int iz = (int)(((p0.z * b01.x + p1.z * b01.y) + p2.z * b23.x) + p3.z * b23.y);
    if ((unsigned)ix < (unsigned)a.X && (unsigned)iy < (unsigned)a.Y && (unsigned)iz < (unsigned)a.Z) return brick_index(ix, iy, iz, a.Sy, a.Sz);
    return 0u;
}

// meta word of a tet record: label 0-7 | level L 8-11 | active 12 | lane in tile 13-17
// pair control words (one int4 per pair, made lane-parallel at the end of phase A):
//   x, y  brick index of the two origins minus the folded magic bits
//   z     rowA 0-11 | rowB 12-23 | labelA 24-31          (first table row of each tet's level)
//   w     nfull 0-4 | nall 5-9 | eps 10 | same level 11 | B active 12 | LA 13-16 | LB 17-20 | labelB 21-28 | same label as the previous pair 29
//
// Voxel widening: a non-negative finite float with bits b is the double with words (b >> 3, b << 29) times 2^896 (zeros
// and denormals included), so sums are kept scaled by 2^-896 -- two shifts per voxel, no conversion instruction, no
// special cases -- and the tet weight v/n^3 carries the factor 2^896 (an exact scaling).  Volumes that hold negative
// values, infinities or NaNs (flag raised when the bricked copy is built) use the general routine and no scaling.
template <typename VoxT, bool SIMPLE, bool WANT_SQ, bool PER_TET>
__global__ void __launch_bounds__(256, G32_MIN_BLOCKS) tet_sums_g32_kernel(const TetG32Args a)
{
    extern __shared__ __align__(16) unsigned char g32_smem[];
    float4 *s_lat = reinterpret_cast<float4 *>(g32_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = g32_smem + kG32Rows * 16 + warp * kG32WarpBytes;
    ulonglong2 *w_rec = reinterpret_cast<ulonglong2 *>(wb);        // [8 field pairs][16 tet pairs]: (f A, f B, f+1 A, f+1 B)
    double *w_w = reinterpret_cast<double *>(wb + 2048);           // [32 slots] v / n^3 (times the widening scale, over the integer denominator)
    int4 *w_ctl = reinterpret_cast<int4 *>(wb + 2304);             // [16 pairs] control words
    uint4 *w_nd = reinterpret_cast<uint4 *>(wb + 2560);            // [32 slots] node ids
    double *w_v = reinterpret_cast<double *>(wb + 3072);           // [32 slots] volume (per-tet outputs only)
    int *w_meta = reinterpret_cast<int *>(wb + 3328);              // [32 slots] meta
    unsigned *w_base = reinterpret_cast<unsigned *>(wb + 3456);    // [32 slots] origin index
    double *w_sw = reinterpret_cast<double *>(wb + 3584);          // slow tets (stored from the end): weight
    int *w_smeta = reinterpret_cast<int *>(wb + 3840);             // ... meta
    __shared__ double sm[8][3 * kMaxPhase];
    for (int i = threadIdx.x; i < kG32Rows; i += blockDim.x) s_lat[i] = a.lat32[i];
    __syncthreads();

    constexpr bool F32S = sizeof(VoxT) == 4 && !VoxTraits<VoxT>::integral && SIMPLE; // scaled widening in use
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.brick);
    const unsigned Sy = a.Sy, Sz = a.Sz;
    const unsigned Ky = Sy - 8u, Kz = Sz - 32u;
    const float MAGIC = 12582912.0f;                 // 1.5 * 2^23: ulp 1
    const unsigned MB = 0x4B400000u, MB2 = 0x4BC00000u, MB4 = 0x4C400000u; // bits of 1.5*2^23, 1.5*2^24, 1.5*2^25
    const unsigned fold = MB * 13u + MB4 * (28u + Kz) + MB2 * Ky;
    const f32x2 M2 = pk2(MAGIC, MAGIC), NM2 = pk2(-MAGIC, -MAGIC), NEG1 = pk2(-1.0f, -1.0f);
    const f32x2 K4 = pk2(0.99999988079071044921875f, 0.99999988079071044921875f);   // 1 - 2^-23
    const f32x2 C4 = pk2(37748736.0f, 37748736.0f);                                 // 1.5*2^25 - 1.5*2^23
    const f32x2 K2 = pk2(0.999999940395355224609375f, 0.999999940395355224609375f); // 1 - 2^-24
    const double dn = VoxTraits<VoxT>::denom();
    const double UNSCALE = F32S ? 0x1p+896 : 1.0, INV_UNSCALE = F32S ? 0x1p-896 : 1.0;
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned n_work = a.tet_end - a.tet_begin;
    const unsigned n_tiles = (n_work + 31) / 32;
    const unsigned total_warps = gridDim.x * 8;

    double acc_tot = 0, acc_sq = 0, acc_vol = 0; // per-label sums of label == lane + 1
    double cur_acc = 0, cur_sq = 0;              // this lane's running sums for label cur_label
    int cur_label = 0;
    unsigned long long nsamp = 0;

    auto flush = [&]() {
        if (cur_label) {
            const double r = warp_sum_d(cur_acc);
            double rs = 0;
            if (WANT_SQ) rs = warp_sum_d(cur_sq);
            if (cur_label <= a.nb_phase && lane == cur_label - 1) { acc_tot += r; acc_sq += rs; }
            cur_acc = 0;
            cur_sq = 0;
        }
    };
    // adds one tet's contribution (this lane's partial sum times the tet weight) to the running sum of its label
    auto contribute = [&](int label, double x, double xs) {
        if (label != cur_label) { flush(); cur_label = label; }
        cur_acc += x;
        if (WANT_SQ) cur_sq += xs;
    };

    // ---- the pair in flight: its gathers have been issued, its sums are not final yet ----
    // Accumulation is deferred: the voxels gathered by one step are added after the NEXT step has computed its indices and
    // just before it issues its own gathers, so the index arithmetic of a step overlaps the memory latency of the previous one.
    double sA = 0, sB = 0, sqA = 0, sqB = 0;   // sums of the open pair (float / double voxels)
    unsigned isA = 0, isB = 0;                 // ... integer voxels
    unsigned long long iqA = 0, iqB = 0;
    VoxT pendA[2], pendB[2];
    pendA[0] = pendA[1] = pendB[0] = pendB[1] = (VoxT)0;
    int pend_n = 0;
    bool open_pair = false, close_pending = false;
    double cl_wA = 0, cl_wB = 0;               // of the pair waiting to be closed
    int cl_labA = 0, cl_labB = 0;
    bool cl_fast = false;

    auto add_voxel = [&](VoxT val, double &s, double &sq, unsigned &is, unsigned long long &iq) {
        if constexpr (VoxTraits<VoxT>::integral) {
            const unsigned q = (unsigned)val;
            is += q;
            if (WANT_SQ) iq += (unsigned long long)q * q;
        } else {
            double I;
            if constexpr (sizeof(VoxT) == 4) {
                if (SIMPLE) {
                    const unsigned b = __float_as_uint((float)val);
                    I = __hiloint2double((int)(b >> 3), (int)(b << 29)); // val * 2^-896, exact
                } else {
                    I = widen_exact((float)val);
                }
            } else {
                I = (double)val;
            }
            s += I;
            if (WANT_SQ) {
                const double Iu = I * UNSCALE; // exact
                sq = fma(Iu, Iu, sq); // second moment: tolerance quantity, fused on purpose
            }
        }
    };
    // adds the gathered voxels of the previous step; closes the previous pair if that step was its last
    auto drain = [&]() {
        if (pend_n) {
            add_voxel(pendA[0], sA, sqA, isA, iqA);
            add_voxel(pendB[0], sB, sqB, isB, iqB);
            if (pend_n == 2) {
                add_voxel(pendA[1], sA, sqA, isA, iqA);
                add_voxel(pendB[1], sB, sqB, isB, iqB);
            }
            pend_n = 0;
        }
        if (close_pending) {
            if constexpr (VoxTraits<VoxT>::integral) {
                sA = (double)isA; sB = (double)isB;
                if (WANT_SQ) { sqA = (double)iqA / dn; sqB = (double)iqB / dn; } // the weight carries the other 1/dn
            }
            double xsA = 0, xsB = 0;
            if (WANT_SQ) { xsA = sqA * (cl_wA * INV_UNSCALE); xsB = sqB * (cl_wB * INV_UNSCALE); } // the weight carries 2^896
            if (cl_fast) { // both labels are the running label
                cur_acc += sA * cl_wA + sB * cl_wB;
                if (WANT_SQ) cur_sq += xsA + xsB;
            } else {
                contribute(cl_labA, sA * cl_wA, xsA);
                if (cl_labB) contribute(cl_labB, sB * cl_wB, xsB);
            }
            sA = sB = sqA = sqB = 0;
            isA = isB = 0;
            iqA = iqB = 0;
            close_pending = false;
        }
    };

    for (unsigned tile = blockIdx.x * 8 + warp; tile < n_tiles; tile += total_warps) {
        const unsigned base_t = a.tet_begin + tile * 32;
        const int nb = (int)min(32u, a.tet_end - base_t);
        if (PER_TET) drain();
        __syncwarp();
        // ---------------- phase A: lane <-> tet ----------------
        int nfast = 0, nslow = 0;
        {
            int label = 0, L = 0, n3 = 0, meta = 0;
            unsigned base = 0;
            bool act = false, fast = false, big = false;
            float f[16];
#pragma unroll
            for (int k = 0; k < 16; k++) f[k] = 0.f;
            double v = 0, w = 0;
            uint4 nd = make_uint4(0, 0, 0, 0);
            if (lane < nb) {
                const unsigned t = base_t + lane;
                label = __ldg(a.mesh.tet_label + t);
                act = label != 0;
                if (act) {
                    nd = __ldg(a.mesh.tet_nodes + t);
                    const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                    const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
                    v = tet_volume(p0, p1, p2, p3);
                    L = tet_level(v);
                    n3 = (L + 1) * (L + 1) * (L + 1);
                    nsamp += (unsigned long long)n3;
                    big = n3 > 32;
                    w = v / (double)n3;
                    if (VoxTraits<VoxT>::integral) w = w / dn;
                    w = w * UNSCALE;
                    const double lox = fmin(fmin(p0.x, p1.x), fmin(p2.x, p3.x)), hix = fmax(fmax(p0.x, p1.x), fmax(p2.x, p3.x));
                    const double loy = fmin(fmin(p0.y, p1.y), fmin(p2.y, p3.y)), hiy = fmax(fmax(p0.y, p1.y), fmax(p2.y, p3.y));
                    const double loz = fmin(fmin(p0.z, p1.z), fmin(p2.z, p3.z)), hiz = fmax(fmax(p0.z, p1.z), fmax(p2.z, p3.z));
                    // the whole tet inside the padded copy, with a margin of two voxels
                    const bool inside = lox >= -6.0 && loy >= -6.0 && loz >= -6.0 && hix < (double)a.X + 6.0 && hiy < (double)a.Y + 6.0 && hiz < (double)a.Z + 6.0;
                    meta = ((label >= 1 && label < 255) ? label : 255) | (L << 8) | (1 << 12) | (lane << 13);
                    if (inside) {
                        // origin: floor(min corner) rounded down to a brick corner (the padding is a multiple of the brick)
                        const int oxi = (((int)floor(lox) + kBrickPad) & ~3) - kBrickPad;
                        const int oyi = (((int)floor(loy) + kBrickPad) & ~1) - kBrickPad;
                        const int ozi = (((int)floor(loz) + kBrickPad) & ~3) - kBrickPad;
                        const double ox = (double)oxi, oy = (double)oyi, oz = (double)ozi;
                        const double rmax = fmax(fmax(hix - ox, hiy - oy), hiz - oz);
                        fast = rmax < 100.0;
                        if (fast) {
                            const double inv4n = 1.0 / (double)(4 * (L + 1));
                            f[0] = (float)((p0.x - ox) - 0.5); f[1] = (float)((p0.y - oy) - 0.5); f[2] = (float)((p0.z - oz) - 0.5);
                            f[3] = 0.5f - (48.0f * (float)rmax + 8.0f) * 5.9604644775390625e-08f;
                            f[4] = (float)((p1.x - p0.x) * inv4n); f[5] = (float)((p1.y - p0.y) * inv4n); f[6] = (float)((p1.z - p0.z) * inv4n);
                            f[7] = (float)((p2.x - p0.x) * inv4n); f[8] = (float)((p2.y - p0.y) * inv4n); f[9] = (float)((p2.z - p0.z) * inv4n);
                            f[10] = (float)((p3.x - p0.x) * inv4n); f[11] = (float)((p3.y - p0.y) * inv4n); f[12] = (float)((p3.z - p0.z) * inv4n);
                            f[13] = (float)ox; f[14] = (float)oy; f[15] = (float)oz;
                            base = brick_index(oxi, oyi, ozi, Sy, Sz) - fold;
                        }
                    }
                } else if (PER_TET) {
                    if (a.tet_total) a.tet_total[t] = -1.0;
                    if (a.tet_vol) a.tet_vol[t] = -1.0;
                    if (a.tet_mean) a.tet_mean[t] = -1.0;
                }
            }
            // order the fast tets by label (first occurrence), tets of at most 32 samples first; sum volumes per label
            int slot = -1;
            unsigned todo = __ballot_sync(full, act);
            while (todo) {
                const int src = __ffs(todo) - 1;
                const int lab = __shfl_sync(full, label, src);
                const bool mine = act && label == lab;
                const unsigned m = __ballot_sync(full, mine);
                const unsigned ms = __ballot_sync(full, mine && fast && !big);
                const unsigned mbg = __ballot_sync(full, mine && fast && big);
                if (mine && fast) slot = nfast + (big ? __popc(ms) + __popc(mbg & lt) : __popc(ms & lt));
                nfast += __popc(ms) + __popc(mbg);
                if (lab >= 1 && lab <= a.nb_phase) {
                    const double vs = warp_sum_d(mine ? v : 0.0);
                    if (lane == lab - 1) acc_vol += vs;
                }
                todo &= ~m;
            }
            const unsigned m_slow = __ballot_sync(full, act && !fast);
            nslow = __popc(m_slow);
            if (act && fast) {
                // field pair i = k / 2 of tet pair p sits at w_rec[i * 16 + p] as (f_k A, f_k B, f_k+1 A, f_k+1 B):
                // consecutive slots write consecutive words (2-way conflicts at worst), a pair reads 8 broadcast LDS.128
                float *recf = reinterpret_cast<float *>(w_rec);
                const int p = slot >> 1, h = slot & 1;
#pragma unroll
                for (int k = 0; k < 16; k++) recf[((k >> 1) * 16 + p) * 4 + (k & 1) * 2 + h] = f[k];
                w_base[slot] = base;
                w_meta[slot] = meta;
                w_w[slot] = w;
                w_nd[slot] = nd;
                if (PER_TET) w_v[slot] = v;
                if (slot == nfast - 1 && h == 0) { // odd count: the last pair's second half mirrors the first, inactive
#pragma unroll
                    for (int k = 0; k < 16; k++) recf[((k >> 1) * 16 + p) * 4 + (k & 1) * 2 + 1] = f[k];
                    w_base[slot + 1] = base;
                    w_meta[slot + 1] = meta & ~(1 << 12);
                    w_w[slot + 1] = 0.0;
                }
            }
            if (act && !fast) { // node ids and volumes share the slot arrays (nfast + nslow <= 32); weight and meta have their own
                const int ss = 31 - __popc(m_slow & lt);
                w_smeta[ss] = meta;
                w_sw[ss] = w * INV_UNSCALE;
                w_nd[ss] = nd;
                if (PER_TET) w_v[ss] = v;
            }
        }
        // pair control words, lane <-> pair
        const int npairs = (nfast + 1) >> 1;
        __syncwarp();
        {
            int4 c = make_int4(0, 0, 0, 0);
            int labA = 0, labBe = 0;
            if (lane < npairs) {
                const int mA = w_meta[2 * lane], mB = w_meta[2 * lane + 1];
                const int LA = (mA >> 8) & 15, LB = (mB >> 8) & 15;
                const bool actB = (mB >> 12) & 1;
                const int n3A = (LA + 1) * (LA + 1) * (LA + 1), n3B = actB ? (LB + 1) * (LB + 1) * (LB + 1) : 0;
                const int rowA = g32_slot_off(LA) * 32, rowB = g32_slot_off(LB) * 32;
                const int nfull = min(n3A, n3B) >> 5, nall = (max(n3A, n3B) + 31) >> 5;
                const int eps = ((0x164 >> LA) | (0x164 >> LB)) & 1; // rows of n = 3, 6, 7, 9 do not sum to exactly one
                labA = mA & 0xff;
                const int labB = actB ? (mB & 0xff) : 0;
                labBe = actB ? labB : labA;
                c.x = (int)w_base[2 * lane];
                c.y = (int)w_base[2 * lane + 1];
                c.z = rowA | (rowB << 12) | (labA << 24);
                c.w = nfull | (nall << 5) | (eps << 10) | ((rowA == rowB ? 1 : 0) << 11) | ((actB ? 1 : 0) << 12) | (LA << 13) | (LB << 17) | (labB << 21);
            }
            const int prevB = __shfl_up_sync(full, labBe, 1);
            if (lane < npairs) {
                if (lane > 0 && labA == labBe && labA == prevB) c.w |= 1 << 29; // the running label is already this pair's
                w_ctl[lane] = c;
            }
        }
        __syncwarp();
        // ---------------- phase B: the warp walks pairs of tets; lane <-> table row, the two halves of a packed register <-> the two tets ----------------
        for (int p = 0; p < npairs; p++) {
            const ulonglong2 r0 = w_rec[0 * 16 + p], r1 = w_rec[1 * 16 + p], r2 = w_rec[2 * 16 + p], r3 = w_rec[3 * 16 + p];
            const ulonglong2 r4 = w_rec[4 * 16 + p], r5 = w_rec[5 * 16 + p], r6 = w_rec[6 * 16 + p], r7 = w_rec[7 * 16 + p];
            const f32x2 q0x = r0.x, q0y = r0.y, q0z = r1.x, thr2 = r1.y, e1x = r2.x, e1y = r2.y, e1z = r3.x, e2x = r3.y;
            const f32x2 e2y = r4.x, e2z = r4.y, e3x = r5.x, e3y = r5.y, e3z = r6.x, o_x = r6.y, o_y = r7.x, o_z = r7.y;
            const int4 ctl = w_ctl[p];
            const unsigned baseA = (unsigned)ctl.x, baseB = (unsigned)ctl.y;
            const int rowA = (ctl.z & 0xfff) + lane, rowB = ((ctl.z >> 12) & 0xfff) + lane;
            const int nfull = ctl.w & 31, nall = (ctl.w >> 5) & 31;
            const int LA = (ctl.w >> 13) & 15, LB = (ctl.w >> 17) & 15;
            const bool actB = (ctl.w >> 12) & 1;
            const int n3A = (LA + 1) * (LA + 1) * (LA + 1), n3B = actB ? (LB + 1) * (LB + 1) * (LB + 1) : 0;
            float thrA, thrB;
            upk2(thr2, thrA, thrB);

            // NS consecutive 32-row slots starting at s; TAIL: rows past the end of either tet are masked
            // SAME: both tets on the same level, one table row serves both halves (broadcast operand, no second LDS)
            auto body = [&](auto ns_c, auto tail_c, auto eps_c, auto same_c, int s) {
                constexpr int NS = decltype(ns_c)::value;
                constexpr bool TAIL = decltype(tail_c)::value;
                constexpr bool EPS = decltype(eps_c)::value;
                constexpr bool SAME = decltype(same_c)::value;
                unsigned idxA[NS], idxB[NS];
                bool unsafeA[NS], unsafeB[NS];
                bool any_unsafe = false;
#pragma unroll
                for (int u = 0; u < NS; u++) {
                    const int ra_i = rowA + (s + u) * 32, rb_i = rowB + (s + u) * 32;
                    const float4 ra = s_lat[ra_i];
                    float4 rb = ra;
                    if (!SAME) rb = s_lat[rb_i];
                    const f32x2 a1 = pk2(ra.x, rb.x), a2 = pk2(ra.y, rb.y), a3 = pk2(ra.z, rb.z);
                    f32x2 qx = q0x, qy = q0y, qz = q0z;
                    if (EPS) {
                        const f32x2 ep = pk2(ra.w, rb.w);
                        qx = fma2(ep, o_x, qx); qy = fma2(ep, o_y, qy); qz = fma2(ep, o_z, qz);
                    }
                    qx = fma2(a1, e1x, qx); qy = fma2(a1, e1y, qy); qz = fma2(a1, e1z, qz);
                    qx = fma2(a2, e2x, qx); qy = fma2(a2, e2y, qy); qz = fma2(a2, e2z, qz);
                    qx = fma2(a3, e3x, qx); qy = fma2(a3, e3y, qy); qz = fma2(a3, e3z, qz);
                    const f32x2 tx = add2(qx, M2), ty = add2(qy, M2), tz = add2(qz, M2);
                    const f32x2 ex = fma2(add2(tx, NM2), NEG1, qx), ey = fma2(add2(ty, NM2), NEG1, qy), ez = fma2(add2(tz, NM2), NEG1, qz);
                    const f32x2 hx = fma2(tx, K4, C4), hy = fma2(ty, K2, M2), hz = fma2(tz, K4, C4);
                    float exA, exB, eyA, eyB, ezA, ezB;
                    upk2(ex, exA, exB); upk2(ey, eyA, eyB); upk2(ez, ezA, ezB);
                    float txA, txB, tyA, tyB, tzA, tzB, hxA, hxB, hyA, hyB, hzA, hzB;
                    upk2(tx, txA, txB); upk2(ty, tyA, tyB); upk2(tz, tzA, tzB);
                    upk2(hx, hxA, hxB); upk2(hy, hyA, hyB); upk2(hz, hzA, hzB);
                    idxA[u] = __float_as_uint(txA) + 28u * __float_as_uint(hxA) + 4u * __float_as_uint(tyA) + Ky * __float_as_uint(hyA) +
                              8u * __float_as_uint(tzA) + Kz * __float_as_uint(hzA) + baseA;
                    idxB[u] = __float_as_uint(txB) + 28u * __float_as_uint(hxB) + 4u * __float_as_uint(tyB) + Ky * __float_as_uint(hyB) +
                              8u * __float_as_uint(tzB) + Kz * __float_as_uint(hzB) + baseB;
                    unsafeA[u] = fmaxf(fmaxf(fabsf(exA), fabsf(eyA)), fabsf(ezA)) >= thrA;
                    unsafeB[u] = fmaxf(fmaxf(fabsf(exB), fabsf(eyB)), fabsf(ezB)) >= thrB;
                    if (TAIL) {
                        unsafeA[u] = unsafeA[u] && (s + u) * 32 + lane < n3A;
                        unsafeB[u] = unsafeB[u] && (s + u) * 32 + lane < n3B;
                    }
                    any_unsafe = any_unsafe || unsafeA[u] || unsafeB[u];
                }
                if (any_unsafe) { // a coordinate within delta of a voxel boundary: the reference's FP64 arithmetic decides
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        bool need = false;
#pragma unroll
                        for (int u = 0; u < NS; u++) need = need || (h ? unsafeB[u] : unsafeA[u]);
                        if (need) {
                            const uint4 nd = w_nd[2 * p + h];
                            const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                            const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
#pragma unroll
                            for (int u = 0; u < NS; u++)
                                if (h ? unsafeB[u] : unsafeA[u]) {
                                    const unsigned e = g32_exact_index(a, p0, p1, p2, p3, h ? LB : LA, (s + u) * 32 + lane);
                                    if (h) idxB[u] = e; else idxA[u] = e;
                                }
                        }
                    }
                }
                if (TAIL) {
#pragma unroll
                    for (int u = 0; u < NS; u++) {
                        if ((s + u) * 32 + lane >= n3A) idxA[u] = 0u; // element 0 is padding: reads 0
                        if ((s + u) * 32 + lane >= n3B) idxB[u] = 0u;
                    }
                }
                drain(); // the previous step's voxels (and, if that step ended a pair, the pair itself)
#pragma unroll
                for (int u = 0; u < NS; u++) { pendA[u] = __ldg(vol + idxA[u]); pendB[u] = __ldg(vol + idxB[u]); }
                pend_n = NS;
            };
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            auto run = [&](auto eps_c, auto same_c) {
                int s = 0;
                for (; s + 2 <= nfull; s += 2) body(I2{}, std::false_type{}, eps_c, same_c, s);
                for (; s < nfull; s++) body(I1{}, std::false_type{}, eps_c, same_c, s);
                for (; s < nall; s++) body(I1{}, std::true_type{}, eps_c, same_c, s);
            };
            if (ctl.w & (1 << 11)) {
                if (ctl.w & (1 << 10)) run(std::true_type{}, std::true_type{});
                else run(std::false_type{}, std::true_type{});
            } else {
                if (ctl.w & (1 << 10)) run(std::true_type{}, std::false_type{});
                else run(std::false_type{}, std::false_type{});
            }
            // the pair is closed by the next drain(), once its last gathers have landed
            {
                const double2 ww = reinterpret_cast<const double2 *>(w_w)[p];
                cl_wA = ww.x; cl_wB = ww.y;
                cl_labA = (ctl.z >> 24) & 0xff; cl_labB = (ctl.w >> 21) & 0xff;
                cl_fast = (ctl.w >> 29) & 1;
                close_pending = true;
            }
            if (PER_TET) { // per-tet outputs: the exact sum of the tet's voxels, then the reference's own expression
                if (pend_n) {
                    add_voxel(pendA[0], sA, sqA, isA, iqA); add_voxel(pendB[0], sB, sqB, isB, iqB);
                    if (pend_n == 2) { add_voxel(pendA[1], sA, sqA, isA, iqA); add_voxel(pendB[1], sB, sqB, isB, iqB); }
                    pend_n = 0;
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (h == 1 && !actB) break;
                    double st;
                    if constexpr (VoxTraits<VoxT>::integral) st = warp_sum_d((double)(h ? isB : isA)) / dn;
                    else st = warp_sum_d(h ? sB : sA) * UNSCALE;
                    if (lane == 0) {
                        const unsigned t = base_t + ((w_meta[2 * p + h] >> 13) & 31);
                        const double vv = w_v[2 * p + h];
                        const double tot = st * vv / (double)(h ? n3B : n3A); // total = total * v / a.size()  (image3d.cpp:371)
                        if (a.tet_total) a.tet_total[t] = tot;
                        if (a.tet_vol) a.tet_vol[t] = vv;
                        if (a.tet_mean) a.tet_mean[t] = tot / vv;
                    }
                }
                drain();
            }
        }
        // ---------------- slow tets: outside the padded copy or wider than 100 voxels; exact arithmetic, the whole warp on one tet ----------------
        if (nslow) drain();
        for (int k = 0; k < nslow; k++) {
            const int ss = 31 - k;
            const int meta = w_smeta[ss];
            const int L = (meta >> 8) & 15, n3 = (L + 1) * (L + 1) * (L + 1);
            const uint4 nd = w_nd[ss];
            const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
            const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
            double s = 0, sq = 0;
            unsigned is = 0;
            unsigned long long iq = 0;
            for (int i = lane; i < n3; i += 32) {
                const VoxT val = __ldg(vol + g32_exact_index(a, p0, p1, p2, p3, L, i));
                if constexpr (VoxTraits<VoxT>::integral) {
                    is += (unsigned)val;
                    if (WANT_SQ) iq += (unsigned long long)val * (unsigned)val;
                } else {
                    double I;
                    if constexpr (sizeof(VoxT) == 4) I = widen_exact((float)val);
                    else I = (double)val;
                    s += I;
                    if (WANT_SQ) sq = fma(I, I, sq);
                }
            }
            if constexpr (VoxTraits<VoxT>::integral) {
                s = (double)is;
                if (WANT_SQ) sq = (double)iq / dn;
            }
            if (PER_TET) {
                double st = warp_sum_d(s);
                if (VoxTraits<VoxT>::integral) st = st / dn;
                if (lane == 0) {
                    const unsigned t = base_t + ((meta >> 13) & 31);
                    const double vv = w_v[ss];
                    const double tot = st * vv / (double)n3;
                    if (a.tet_total) a.tet_total[t] = tot;
                    if (a.tet_vol) a.tet_vol[t] = vv;
                    if (a.tet_mean) a.tet_mean[t] = tot / vv;
                }
            }
            const double w = w_sw[ss]; // true doubles here
            contribute(meta & 0xff, s * w, WANT_SQ ? sq * w : 0.0);
        }
    }
    drain();
    flush();
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) nsamp += __shfl_xor_sync(full, nsamp, m);
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
    if (lane == 0 && nsamp) atomicAdd(a.sample_count, nsamp);
}

} // namespace dsc
