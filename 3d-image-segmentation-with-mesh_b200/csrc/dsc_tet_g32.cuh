// dsc_tet_g32.cuh -- per-tet integration, warp-per-tet form over a bricked copy of the volume (tet variant bit 1024).
//
// Same arithmetic contract as tet_integrate_w2_kernel (dsc_kernels.cuh): image3d::get_tetra_intensity
// (DEMO/image3d.cpp:351-382) batched over all tets and fused with the per-label reduction; the sample -> voxel
// classification is bit-identical to the reference (FP32 lattice filter with a rigorous guard band, exact FP64 re-evaluation
// inside the band), sums are FP64.  It attacks the two limits the profile of w2 shows (profiles/README.md: the L1 data pipe
// at one wavefront per distinct 128 B line touched by a gather -- 16.8 lines per 32-lane gather on the x-fastest volume --
// and instruction issue at ~52 instructions per sample):
//
//  * bricked volume: gathers go to a copy of the volume laid out in 4x2x4-voxel bricks (32 voxels; 128 B for float), padded
//    by 8 voxels on every side (dsc_kernels.cuh: brick_index).  With all 32 lanes on ONE tet a gather touches 8.0 lines
//    instead of 16.8 (scripts/line_sim.py reproduces both figures on the C4 mesh; ncu confirms).  The padding replaces the
//    bounds test of image3d::get_value, so every tet inside the padded region takes the fast path.
//  * tet-local origins are aligned to brick corners, so the brick index and the offset inside the brick are both linear
//    in r and floor(r / brick): idx = base + rx + 28 hx + 4 ry + (Sy-8) hy + 8 rz + (Sz-32) hz.
//  * packed FP32 (Blackwell FFMA2 / FADD2, PTX fma.rn.f32x2): one lane evaluates two samples per instruction -- the two
//    halves of a pair of tets (pairs are formed in phase A from tets of the same label and sample level, so both halves
//    use the same table row) -- which halves the issue slots of the position arithmetic.  floor() is the 1.5*2^23 trick;
//    floor(r/4) and floor(r/2) come from one more FFMA2 each on the already rounded value: fma(t, 1-2^-23, 4.5*2^23)
//    rounds M4 + r - 1.5 - r 2^-23 at ulp 4 (tests/test_brick_index.py emulates all of this on the host).
//  * no per-tet cross-lane reduction: each lane multiplies its partial sum by the tet's v/n^3 and keeps a running sum for
//    the current label; a warp reduction happens only when the label changes (tets are sorted by label inside a tile).
//
// MEASURED (C4, B200): 0.47 ms against 0.39 ms of w2, so w2 stays the default.  The index arithmetic is down to 24
// instructions per sample (w2: 52) and the L1 wavefronts of the gathers are halved, but the packed pair records cost 126
// registers (16 warps per SM), and with four gathers in flight per warp the kernel is bound by memory latency: 31 % of the
// stall samples sit on the first use of the gathered voxels although the next step's indices are computed between the
// loads and that use.  Register-destination loads cannot be kept in flight any longer than that: ptxas drains every pending
// load at the first branch it meets, warp-uniform or not.  cp.async gathers (4 B per lane) lift that limit but load the L1
// pipe to 79 % (0.54 ms); L2 prefetches of the next step's addresses cost more than they hide (0.79 ms).  DESIGN.md
// section 9 has the table.
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

constexpr int kG32Slots = 66;                     // sum over n = 1..9 of ceil(n^3 / 32)
constexpr int kG32Rows = kG32Slots * 32;          // padded table rows
constexpr int kG32Pairs = 24;                     // tet pairs a tile can hold (32 tets plus the odd member of each label/level class)
constexpr int kG32WarpBytes = 8 * kG32Pairs * 16 + 2 * kG32Pairs * 8 + kG32Pairs * 16 + 512 + 256 + 2 * kG32Pairs * 4 + 2 * kG32Pairs * 4 + 128 + 256;
constexpr int kG32SmemBytes = kG32Rows * 16 + 8 * kG32WarpBytes;

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

// meta word of a tet: label 0-7 | level L 8-11 | active 12 | lane in tile 13-17
// pair control words (one int4 per pair, made lane-parallel at the end of phase A):
//   x, y  brick index of the two origins minus the folded magic bits
//   z     first table row of the level 0-11 | label 12-19 | tile lane of A 20-24 | tile lane of B 25-29
//   w     nfull 0-4 | nall 5-9 | eps 10 | B active 12 | L 13-16 | label equals the previous pair's 29
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
    constexpr int NP = kG32Pairs;
    ulonglong2 *w_rec = reinterpret_cast<ulonglong2 *>(wb);                 // [8 field pairs][NP tet pairs]: (f A, f B, f+1 A, f+1 B)
    double *w_w = reinterpret_cast<double *>(wb + 128 * NP);                // [2 NP slots] v / n^3 (times the widening scale, over the integer denominator)
    int4 *w_ctl = reinterpret_cast<int4 *>(wb + 144 * NP);                  // [NP pairs] control words
    uint4 *w_nd = reinterpret_cast<uint4 *>(wb + 160 * NP);                 // [32 tile lanes] node ids
    double *w_v = reinterpret_cast<double *>(wb + 160 * NP + 512);          // [32 tile lanes] volume (per-tet outputs only)
    int *w_meta = reinterpret_cast<int *>(wb + 160 * NP + 768);             // [2 NP slots] meta
    unsigned *w_base = reinterpret_cast<unsigned *>(wb + 168 * NP + 768);   // [2 NP slots] origin index
    int *w_slow = reinterpret_cast<int *>(wb + 176 * NP + 768);             // [32] meta of the slow tets
    double *w_sw = reinterpret_cast<double *>(wb + 176 * NP + 896);         // [32] their weights (unscaled)
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
        if (__any_sync(full, cur_label != 0)) {
            const double r = warp_sum_d(cur_acc);
            double rs = 0;
            if (WANT_SQ) rs = warp_sum_d(cur_sq);
            if (cur_label <= a.nb_phase && lane == cur_label - 1) { acc_tot += r; acc_sq += rs; }
            cur_acc = 0;
            cur_sq = 0;
        }
    };
    // adds a contribution (this lane's partial sum times the tet weight) to the running sum of its label
    auto contribute = [&](int label, double x, double xs) {
        if (label != cur_label) { flush(); cur_label = label; }
        cur_acc += x;
        if (WANT_SQ) cur_sq += xs;
    };

    // sums of the open pair (this lane's rows): float / double voxels, integer voxels
    double sA = 0, sB = 0, sqA = 0, sqB = 0;
    unsigned isA = 0, isB = 0;
    unsigned long long iqA = 0, iqB = 0;

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

    for (unsigned tile = blockIdx.x * 8 + warp; tile < n_tiles; tile += total_warps) {
        const unsigned base_t = a.tet_begin + tile * 32;
        const int nb = (int)min(32u, a.tet_end - base_t);
        __syncwarp();
        // ---------------- phase A: lane <-> tet ----------------
        int nslots = 0, nslow = 0;
        {
            int label = 0, L = 0, n3 = 0, meta = 0;
            unsigned base = 0;
            bool act = false, fast = false;
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
                    w = v / (double)n3;
                    if (VoxTraits<VoxT>::integral) w = w / dn;
                    const double lox = fmin(fmin(p0.x, p1.x), fmin(p2.x, p3.x)), hix = fmax(fmax(p0.x, p1.x), fmax(p2.x, p3.x));
                    const double loy = fmin(fmin(p0.y, p1.y), fmin(p2.y, p3.y)), hiy = fmax(fmax(p0.y, p1.y), fmax(p2.y, p3.y));
                    const double loz = fmin(fmin(p0.z, p1.z), fmin(p2.z, p3.z)), hiz = fmax(fmax(p0.z, p1.z), fmax(p2.z, p3.z));
                    // the whole tet inside the padded copy, with a margin of two voxels
                    const bool inside = lox >= -6.0 && loy >= -6.0 && loz >= -6.0 && hix < (double)a.X + 6.0 && hiy < (double)a.Y + 6.0 && hiz < (double)a.Z + 6.0;
                    label = (label >= 1 && label < 255) ? label : 255; // labels beyond nb_phase are processed but belong to no phase
                    meta = label | (L << 8) | (1 << 12) | (lane << 13);
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
            // Order the fast tets by label (first occurrence) and, inside a label, by level; every (label, level) class
            // starts on an even slot so that the two tets of a pair always share label and table rows.  Volumes are
            // summed per label on the way.
            int slot = -1;
            bool mirror = false; // this tet is the odd one of its class: it also fills the second half of its pair, inactive
            unsigned todo = __ballot_sync(full, act);
            while (todo) {
                const int lab = __shfl_sync(full, label, __ffs(todo) - 1);
                const bool in_lab = act && label == lab;
                const unsigned m = __ballot_sync(full, in_lab);
                unsigned todo_l = __ballot_sync(full, in_lab && fast);
                while (todo_l) {
                    const int lev = __shfl_sync(full, L, __ffs(todo_l) - 1);
                    const bool mine = in_lab && fast && L == lev;
                    const unsigned mc = __ballot_sync(full, mine);
                    const int cnt = __popc(mc);
                    if (mine) {
                        slot = nslots + __popc(mc & lt);
                        mirror = (cnt & 1) && slot == nslots + cnt - 1;
                    }
                    nslots += (cnt + 1) & ~1;
                    todo_l &= ~mc;
                }
                if (lab <= a.nb_phase) {
                    const double vs = warp_sum_d(in_lab ? v : 0.0);
                    if (lane == lab - 1) acc_vol += vs;
                }
                todo &= ~m;
            }
            if (slot >= 2 * NP) { fast = false; slot = -1; } // more classes than the tile record can hold: the exact path takes the rest
            if (nslots > 2 * NP) nslots = 2 * NP;
            const unsigned m_slow = __ballot_sync(full, act && !fast);
            nslow = __popc(m_slow);
            if (act) {
                w_nd[lane] = nd;
                if (PER_TET) w_v[lane] = v;
            }
            if (act && fast) {
                // field pair i = k / 2 of tet pair p sits at w_rec[i * NP + p] as (f_k A, f_k B, f_k+1 A, f_k+1 B):
                // consecutive slots write consecutive words, a pair reads 8 broadcast LDS.128
                float *recf = reinterpret_cast<float *>(w_rec);
                const int p = slot >> 1, h = slot & 1;
#pragma unroll
                for (int k = 0; k < 16; k++) recf[((k >> 1) * NP + p) * 4 + (k & 1) * 2 + h] = f[k];
                w_base[slot] = base;
                w_meta[slot] = meta;
                w_w[slot] = w * UNSCALE;
                if (mirror) {
#pragma unroll
                    for (int k = 0; k < 16; k++) recf[((k >> 1) * NP + p) * 4 + (k & 1) * 2 + 1] = f[k];
                    w_base[slot + 1] = base;
                    w_meta[slot + 1] = meta & ~(1 << 12);
                    w_w[slot + 1] = 0.0;
                }
            } else if (act) {
                const int ss = __popc(m_slow & lt);
                w_slow[ss] = meta;
                w_sw[ss] = w;
            }
        }
        // pair control words, lane <-> pair
        const int npairs = nslots >> 1;
        __syncwarp();
        {
            int4 c = make_int4(0, 0, 0, 0);
            int lab = 0;
            if (lane < npairs) {
                const int mA = w_meta[2 * lane], mB = w_meta[2 * lane + 1];
                const int L = (mA >> 8) & 15;
                const bool actB = (mB >> 12) & 1;
                const int n3 = (L + 1) * (L + 1) * (L + 1);
                const int nfull = actB ? n3 >> 5 : 0, nall = (n3 + 31) >> 5;
                const int eps = (0x164 >> L) & 1; // rows of n = 3, 6, 7, 9 do not sum to exactly one
                lab = mA & 0xff;
                c.x = (int)w_base[2 * lane];
                c.y = (int)w_base[2 * lane + 1];
                c.z = (g32_slot_off(L) * 32) | (lab << 12) | (((mA >> 13) & 31) << 20) | (((mB >> 13) & 31) << 25);
                c.w = nfull | (nall << 5) | (eps << 10) | ((actB ? 1 : 0) << 12) | (L << 13);
            }
            const int prev = __shfl_up_sync(full, lab, 1);
            if (lane < npairs) {
                if (lane > 0 && lab == prev) c.w |= 1 << 29; // the running label is already this pair's
                w_ctl[lane] = c;
            }
        }
        __syncwarp();
        // ---------------- phase B: the warp walks pairs of tets; lane <-> table row, the two halves of a packed register <-> the two tets ----------------
        // A step is one or two 32-row slots of one pair.  Software pipeline, one step deep: an iteration issues the gathers of
        // the CURRENT step (its indices were computed by the previous iteration), computes the indices of the NEXT step and
        // only then adds the gathered voxels -- all in one straight-line block, because the compiler drains every load in
        // flight at the first branch it meets (ncu: a third of all stalls sat on such a branch).  Everything branchy (loading
        // the next pair's record, the exact re-evaluation inside the guard band, closing a pair) happens between blocks.
        if (npairs > 0) {
#define UNI(c) __any_sync(full, (c))
            f32x2 q0x = 0, q0y = 0, q0z = 0, e1x = 0, e1y = 0, e1z = 0, e2x = 0, e2y = 0, e2z = 0, e3x = 0, e3y = 0, e3z = 0, o_x = 0, o_y = 0, o_z = 0;
            float thrA = 0, thrB = 0;
            unsigned baseA = 0, baseB = 0;
            int row0 = 0, nfull = 0, nall = 0, L = 0, n3 = 0, ctl_z = 0, ctl_w = 0;
            bool actB = false;
            auto load_record = [&](int p) {
                const ulonglong2 r0 = w_rec[0 * NP + p], r1 = w_rec[1 * NP + p], r2 = w_rec[2 * NP + p], r3 = w_rec[3 * NP + p];
                const ulonglong2 r4 = w_rec[4 * NP + p], r5 = w_rec[5 * NP + p], r6 = w_rec[6 * NP + p], r7 = w_rec[7 * NP + p];
                q0x = r0.x; q0y = r0.y; q0z = r1.x; upk2(r1.y, thrA, thrB); e1x = r2.x; e1y = r2.y; e1z = r3.x; e2x = r3.y;
                e2y = r4.x; e2z = r4.y; e3x = r5.x; e3y = r5.y; e3z = r6.x; o_x = r6.y; o_y = r7.x; o_z = r7.y;
                const int4 ctl = w_ctl[p];
                baseA = (unsigned)ctl.x; baseB = (unsigned)ctl.y;
                ctl_z = ctl.z; ctl_w = ctl.w;
                row0 = (ctl.z & 0xfff) + lane;
                nfull = ctl.w & 31; nall = (ctl.w >> 5) & 31;
                L = (ctl.w >> 13) & 15;
                actB = (ctl.w >> 12) & 1;
                n3 = (L + 1) * (L + 1) * (L + 1);
            };
            // one 32-row slot: packed position, floor, guard-band test and brick index of this lane's row for both tets
            auto compute = [&](auto eps_c, int sl, unsigned &iA, unsigned &iB, bool &uA, bool &uB) {
                const float4 r = s_lat[row0 + sl * 32];
                const f32x2 a1 = pk2(r.x, r.x), a2 = pk2(r.y, r.y), a3 = pk2(r.z, r.z);
                f32x2 qx = q0x, qy = q0y, qz = q0z;
                if (decltype(eps_c)::value) {
                    const f32x2 ep = pk2(r.w, r.w);
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
                iA = __float_as_uint(txA) + 28u * __float_as_uint(hxA) + 4u * __float_as_uint(tyA) + Ky * __float_as_uint(hyA) +
                     8u * __float_as_uint(tzA) + Kz * __float_as_uint(hzA) + baseA;
                iB = __float_as_uint(txB) + 28u * __float_as_uint(hxB) + 4u * __float_as_uint(tyB) + Ky * __float_as_uint(hyB) +
                     8u * __float_as_uint(tzB) + Kz * __float_as_uint(hzB) + baseB;
                uA = fmaxf(fmaxf(fabsf(exA), fabsf(eyA)), fabsf(ezA)) >= thrA;
                uB = fmaxf(fmaxf(fabsf(exB), fabsf(eyB)), fabsf(ezB)) >= thrB;
            };
            // indices of the step (record in registers, slot s): A0, B0, A1, B1; element 0 of the copy is padding and reads 0
            unsigned n_i[4];
            bool n_u[4];
            auto compute_step = [&](auto two_c, auto eps_c, int s) {
                n_i[2] = n_i[3] = 0u;
                n_u[2] = n_u[3] = false;
                compute(eps_c, s, n_i[0], n_i[1], n_u[0], n_u[1]);
                if (decltype(two_c)::value) { // two full slots: no masks
                    compute(eps_c, s + 1, n_i[2], n_i[3], n_u[2], n_u[3]);
                } else { // one slot; rows past the end of the level (and the whole second half of an odd pair) are masked
                    const bool in = s * 32 + lane < n3;
                    if (!in) { n_i[0] = 0u; n_u[0] = false; }
                    if (!in || !actB) { n_i[1] = 0u; n_u[1] = false; }
                }
            };
            // the current step
            unsigned c_i[4] = {0u, 0u, 0u, 0u};
            bool c_u[4] = {false, false, false, false};
            int c_L = 0, c_s = 0, c_z = 0, c_w = 0, c_p = 0;
            bool c_last = false;
            // the next step: pair p, slot s, record of pair p in registers
            int p = 0, s = 0;
            load_record(0);
            bool have_next = true;
            {   // prologue: the first step's indices
                const bool two = UNI(2 <= nfull), eps = UNI((ctl_w >> 10) & 1);
                if (two) { if (eps) compute_step(std::true_type{}, std::true_type{}, 0); else compute_step(std::true_type{}, std::false_type{}, 0); }
                else { if (eps) compute_step(std::false_type{}, std::true_type{}, 0); else compute_step(std::false_type{}, std::false_type{}, 0); }
            }
            while (true) {
                // ---- next becomes current; find the step after it and, if it opens a new pair, load that pair's record ----
#pragma unroll
                for (int k = 0; k < 4; k++) { c_i[k] = n_i[k]; c_u[k] = n_u[k]; }
                c_L = L; c_s = s; c_z = ctl_z; c_w = ctl_w; c_p = p;
                s += UNI(s + 2 <= nfull) ? 2 : 1;
                c_last = s >= nall;
                if (UNI(c_last)) {
                    p++;
                    s = 0;
                    have_next = p < npairs;
                    if (UNI(have_next)) load_record(p);
                }
                if (UNI(c_u[0] || c_u[1] || c_u[2] || c_u[3])) { // a coordinate within delta of a voxel boundary: the reference's FP64 arithmetic decides
                    auto fix = [&](int tl, bool u0, bool u1, unsigned &i0, unsigned &i1) {
                        if (!(u0 || u1)) return;
                        const uint4 nd = w_nd[tl];
                        const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y);
                        const V3 p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
                        if (u0) i0 = g32_exact_index(a, p0, p1, p2, p3, c_L, c_s * 32 + lane);
                        if (u1) i1 = g32_exact_index(a, p0, p1, p2, p3, c_L, (c_s + 1) * 32 + lane);
                    };
                    fix((c_z >> 20) & 31, c_u[0], c_u[2], c_i[0], c_i[2]);
                    fix((c_z >> 25) & 31, c_u[1], c_u[3], c_i[1], c_i[3]);
                }
                // ---- straight-line block: gathers of the current step, indices of the next one, then the accumulation ----
                const bool two = UNI(s + 2 <= nfull), eps = UNI((ctl_w >> 10) & 1);
                auto block = [&](auto next_c, auto two_c, auto eps_c) {
                    const VoxT v0 = __ldg(vol + c_i[0]), v1 = __ldg(vol + c_i[1]), v2 = __ldg(vol + c_i[2]), v3 = __ldg(vol + c_i[3]);
                    if (decltype(next_c)::value) compute_step(two_c, eps_c, s);
                    add_voxel(v0, sA, sqA, isA, iqA);
                    add_voxel(v1, sB, sqB, isB, iqB);
                    add_voxel(v2, sA, sqA, isA, iqA);
                    add_voxel(v3, sB, sqB, isB, iqB);
                };
                if (UNI(have_next)) {
                    if (two) { if (eps) block(std::true_type{}, std::true_type{}, std::true_type{}); else block(std::true_type{}, std::true_type{}, std::false_type{}); }
                    else { if (eps) block(std::true_type{}, std::false_type{}, std::true_type{}); else block(std::true_type{}, std::false_type{}, std::false_type{}); }
                } else {
                    block(std::false_type{}, std::false_type{}, std::false_type{});
                }
                // ---- the current step was the last of its pair: close the pair ----
                if (UNI(c_last)) {
                    const int cn3 = (c_L + 1) * (c_L + 1) * (c_L + 1);
                    const bool c_actB = (c_w >> 12) & 1, c_fastlab = (c_w >> 29) & 1;
                    if (PER_TET) { // per-tet outputs: the exact sum of the tet's voxels, then the reference's own expression
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if (h == 1 && !UNI(c_actB)) break;
                            double st;
                            if constexpr (VoxTraits<VoxT>::integral) st = warp_sum_d((double)(h ? isB : isA)) / dn;
                            else st = warp_sum_d(h ? sB : sA) * UNSCALE;
                            if (lane == 0) {
                                const int tl = (c_z >> (20 + 5 * h)) & 31;
                                const unsigned t = base_t + tl;
                                const double vv = w_v[tl];
                                const double tot = st * vv / (double)cn3; // total = total * v / a.size()  (image3d.cpp:371)
                                if (a.tet_total) a.tet_total[t] = tot;
                                if (a.tet_vol) a.tet_vol[t] = vv;
                                if (a.tet_mean) a.tet_mean[t] = tot / vv;
                            }
                        }
                    }
                    const double2 ww = reinterpret_cast<const double2 *>(w_w)[c_p];
                    const int lab = (c_z >> 12) & 0xff;
                    if constexpr (VoxTraits<VoxT>::integral) {
                        sA = (double)isA; sB = (double)isB;
                        if (WANT_SQ) { sqA = (double)iqA / dn; sqB = (double)iqB / dn; } // the weight carries the other 1/dn
                    }
                    if (UNI(!c_fastlab && lab != cur_label)) { flush(); cur_label = lab; }
                    cur_acc += sA * ww.x + sB * ww.y;
                    if (WANT_SQ) cur_sq += sqA * (ww.x * INV_UNSCALE) + sqB * (ww.y * INV_UNSCALE); // the weight carries 2^896
                    sA = sB = sqA = sqB = 0;
                    isA = isB = 0;
                    iqA = iqB = 0;
                }
                if (!UNI(have_next)) break;
            }
#undef UNI
        }
        // ---------------- slow tets: outside the padded copy or wider than 100 voxels; exact arithmetic, the whole warp on one tet ----------------
        for (int k = 0; k < nslow; k++) {
            const int meta = w_slow[k];
            const int L = (meta >> 8) & 15, n3 = (L + 1) * (L + 1) * (L + 1);
            const int tl = (meta >> 13) & 31;
            const uint4 nd = w_nd[tl];
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
                    const unsigned t = base_t + tl;
                    const double vv = w_v[tl];
                    const double tot = st * vv / (double)n3;
                    if (a.tet_total) a.tet_total[t] = tot;
                    if (a.tet_vol) a.tet_vol[t] = vv;
                    if (a.tet_mean) a.tet_mean[t] = tot / vv;
                }
            }
            const double w = w_sw[k]; // true doubles here
            contribute(meta & 0xff, s * w, WANT_SQ ? sq * w : 0.0);
        }
    }
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
