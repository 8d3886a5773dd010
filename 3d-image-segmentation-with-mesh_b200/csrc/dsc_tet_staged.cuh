// dsc_tet_staged.cuh -- per-tet voxel integration with the tile's voxel box staged in shared memory by TMA.
//
// image3d::get_tetra_intensity (DEMO/image3d.cpp:351-382) over all tets, fused with the per-label reduction, as the hot path of
// the image-energy evaluation.  What is different from tet_integrate_w2_kernel (dsc_kernels.cuh):
//
//  * lane <-> tet for the whole tile of 32 consecutive tets.  The sample table of the reference is a lattice
//    (tools/gen_tables.py restates its generator, DEMO/image3d.cpp:256-306): the unit cube is cut into n^3 cells, every cell
//    into six tets, and the centroids with x+y+z < 1 are kept.  In barycentric numerators a_k = 4n b_k (exact up to the
//    6-decimal rounding eta_k of the printed literals) a row is
//         a1 = 4j + oy,  a2 = 4i + ox,  a3 = 4k + oz,   (ox,oy,oz) one of the six kOff vectors, (i,j,k) the cell,
//    so with D_k = (P_k - P0)/n and E_k = D_k/4 the sample position is
//         P0 + j D1 + i D2 + k D3  +  (oy E1 + ox E2 + oz E3)   (+ eps P0 + tau, see below).
//    With the row index a compile-time constant (every lane of the warp is at the same row of its own tet) the cell corner is
//    shared by up to six samples and a sample costs ONE packed add for x,y and half a packed add for z, instead of twelve FMAs
//    and a table row from shared memory.  The rows of level n-1 are a subset of the (cell, offset) slots of level n
//    (scripts/lattice_structure.py), so tets of level NB-1 ride the level-NB instruction stream with their own D, E.
//  * the voxels a tile can touch form a small box (32 consecutive tets of a DSC grid mesh: ~52 x 10 x 10 voxels).  One lane
//    per z slice asks the TMA unit for it (cp.async.bulk.tensor.3d, one tensor map per (row bytes, rows) pair), the warp waits
//    on its mbarrier and every gather is an LDS: no L1 tags, no sectors, no DRAM latency inside the sample loop.
//  * one CTA per SM: eight producer warps (tile claim, ids, positions, FP64 geometry, box, TMA) and eight consumer warps (the sample stream)
//    around a ring of box slots in shared memory; tiles are claimed dynamically in chunks of 1-4.
//  * the level-4 stream is three loops over its outer cells (constant tables), not straight-line code: fully unrolled it is 17 KB, and with
//    the producer's 17 KB the hot code missed the SM's instruction cache (hit rate 83 %, a quarter of the stream's stall samples).
//  * per-label sums are accumulated as 64-bit fixed point (exact integer adds of per-tet contributions rounded to 2^-k), so
//    the result does not depend on which warp took which tile, nor on the number of GPUs the tets are partitioned over.
//
// FP32 filter (what feeds an integer decision must equal the reference's FP64 result): with O the integer
// origin of the tet, q = (P0 - O - 1/2) + lattice terms evaluated in float, t = q + 1.5*2^23 holds round(q) = floor(sample - O)
// in its mantissa unless q is within delta of a half-integer.  delta bounds |q_float - q_reference| (derivation at stg_delta);
// a tet with any |q - round(q)| >= 1/2 - delta is recomputed with the reference's exact FP64 sequence (tet_exact_list_kernel), as are
// tets of level n >= 5, tets reaching outside the volume, and tets whose box does not fit a slot.
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

#ifndef STG_PROD
#define STG_PROD 8
#endif
#ifndef STG_CONS
#define STG_CONS 8
#endif
#ifndef STG_UNROLL_S2
#define STG_UNROLL_S2 2 // cells of the i + j + k = 2 loop of the level-4 stream per iteration
#endif
#ifndef STG_PROD_SLEEP
#define STG_PROD_SLEEP 100 // ns between two looks of a producer at its slot's release counter
#endif
constexpr int kStgProd = STG_PROD, kStgCons = STG_CONS; // producer / consumer warps of a CTA
constexpr int kStgMaxSlots = 16;
constexpr int kStgMaxChunk = 4; // most tiles a producer claims with one atomic
constexpr int kStgMaxUnits = 16; // tensor-map family: boxes of 1..16 units of 16 bytes by 1..16 rows by 1..16 slices
constexpr int kStgMaxRows = 16;
constexpr int kStgMaxSlices = 16;

// the six centroid offsets (ox, oy, oz) of the reference's cube split, in table order (DEMO/image3d.cpp:274-291)
__host__ __device__ constexpr int stg_off(int t, int c)
{
    constexpr int o[6][3] = {{1, 1, 1}, {2, 1, 2}, {1, 2, 3}, {3, 2, 1}, {2, 3, 2}, {3, 3, 3}};
    return o[t][c];
}
// lowest level whose table holds slot (cell sum s = i+j+k, offset t): the outermost layer of cells keeps one tet, the next five
__host__ __device__ constexpr int stg_need(int s, int t) { return s + (t == 0 ? 1 : (t <= 4 ? 2 : 3)); }
// sign of eps = sum_k eta_k for a level-3 row, in units of 1e-6: a_k = 1 mod 3 was printed rounded down, 2 mod 3 rounded up
__host__ __device__ constexpr int stg_eps3(int i, int j, int k, int t)
{
    const int a1 = 4 * j + stg_off(t, 1), a2 = 4 * i + stg_off(t, 0), a3 = 4 * k + stg_off(t, 2), a0 = 12 - a1 - a2 - a3;
    int c = 0;
    const int a[4] = {a0, a1, a2, a3};
    for (int m = 0; m < 4; m++) c += (a[m] % 3 == 2) - (a[m] % 3 == 1);
    return c / 3; // -1, 0 or +1 (the sum of the four is 0 mod 3)
}

DSC_D unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
DSC_D f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
DSC_D float fmax3abs(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));
    return r;
}
template <typename VoxT> DSC_D VoxT lds_vox(unsigned addr);
template <> DSC_D float lds_vox<float>(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <> DSC_D unsigned short lds_vox<unsigned short>(unsigned addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
template <> DSC_D unsigned char lds_vox<unsigned char>(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return (unsigned char)v;
}

// per-lane state of the block path
struct StgLane {
    f32x2 q0xy, q0zz;                               // P0 - O - 1/2 (z in both halves)
    f32x2 d1xy, d1zz, d2xy, d2zz, d3xy, d3zz;       // (P_k - P0)/n
    f32x2 oxy[6], ozz[3];                           // centroid offsets; z of offsets (0,1) (2,3) (4,5) paired
    f32x2 epxy, epzz;                               // 1e-6 * P0 (level 3 only: the 6-decimal literals of n = 3 do not sum to 1)
    unsigned base;                                  // shared address of voxel O in the box, minus the folded 1.5*2^23 bit patterns
    unsigned bx, bxy;                               // byte pitch of a box row / slice
    unsigned zero_addr;                             // shared address of a zero word (slots of level NB a level NB-1 tet does not have)
    float thr;                                      // 1/2 - delta
};

// per-tet sums of the sampled voxels: four interleaved partial sums (the adds of consecutive samples do not wait for each other;
// float voxel sums of this length are exact in double, integer sums are exact anyway, so the split changes nothing)
template <typename VoxT, bool WANT_SQ, bool F32S> struct StgAcc {
    double s[4] = {0, 0, 0, 0}, sq[2] = {0, 0};
    unsigned isum[2] = {0, 0};
    unsigned long long isq = 0;
    float emax = 0.f;
    DSC_D void add(VoxT val, int c)
    {
        if constexpr (VoxTraits<VoxT>::integral) {
            const unsigned q = (unsigned)val;
            isum[c & 1] += q;
            if (WANT_SQ) isq += (unsigned long long)q * q;
        } else {
            double I;
            if constexpr (F32S) I = widen_scaled((float)val);
            else if constexpr (sizeof(VoxT) == 4) I = widen_exact((float)val);
            else I = (double)val;
            s[c & 3] += I;
            if (WANT_SQ) {
                const double Iu = F32S ? I * 0x1p+896 : I;   // exact
                sq[c & 1] = fma(Iu, Iu, sq[c & 1]);          // second moment: tolerance quantity, fused on purpose
            }
        }
    }
    // the true sums
    DSC_D void result(double &sum, double &sumsq) const
    {
        if constexpr (VoxTraits<VoxT>::integral) {
            const double dn = VoxTraits<VoxT>::denom();
            sum = (double)(isum[0] + isum[1]) / dn;
            sumsq = WANT_SQ ? (double)isq / (dn * dn) : 0.0;
        } else {
            const double t = (s[0] + s[1]) + (s[2] + s[3]);
            sum = F32S ? t * 0x1p+896 : t;
            sumsq = sq[0] + sq[1];
        }
    }
};

// one sample: position bits -> shared address -> voxel
template <typename VoxT, bool SEL> DSC_D unsigned stg_addr(const StgLane &r, float tx, float ty, float tz, bool full)
{
    constexpr int SH = sizeof(VoxT) == 4 ? 2 : (sizeof(VoxT) == 2 ? 1 : 0);
    const unsigned addr = r.base + (__float_as_uint(tx) << SH) + __float_as_uint(ty) * r.bx + __float_as_uint(tz) * r.bxy;
    return SEL ? (full ? addr : r.zero_addr) : addr;
}

// One or two samples of a cell, positions qa, qb (x, y) and qzz (z of both): voxel index bits by the magic-number add, guard-band residue,
// shared address, load, add.  sel_a / sel_b: the slot exists at level NB only (a level NB-1 lane reads the zero word instead).
// (two, sel_a, sel_b, cnt are constants at every call once the callers are unrolled / inlined)
template <typename VoxT, bool WANT_SQ, bool F32S>
DSC_D void stg_pair(const StgLane &r, bool full, StgAcc<VoxT, WANT_SQ, F32S> &acc, f32x2 qa, f32x2 qb, f32x2 qzz, bool two, bool sel_a, bool sel_b, int cnt, float &em,
                    float &em1)
{
    const float MAGIC = 12582912.0f; // 1.5 * 2^23
    const f32x2 M2 = pk2(MAGIC, MAGIC);
    const f32x2 ta = add2(qa, M2), tz2 = add2(qzz, M2);
    const f32x2 ea2 = sub2(qa, sub2(ta, M2)), ez2 = sub2(qzz, sub2(tz2, M2));
    float tax, tay, tz0, tz1, eax, eay, ez0, ez1;
    upk2(ta, tax, tay); upk2(tz2, tz0, tz1); upk2(ea2, eax, eay); upk2(ez2, ez0, ez1);
    em = fmax3abs(em, eax, eay);
    const unsigned addr_a = sel_a ? stg_addr<VoxT, true>(r, tax, tay, tz0, full) : stg_addr<VoxT, false>(r, tax, tay, tz0, full);
    const VoxT va = lds_vox<VoxT>(addr_a);
    if (two) {
        const f32x2 tb = add2(qb, M2);
        const f32x2 eb2 = sub2(qb, sub2(tb, M2));
        float tbx, tby, ebx, eby;
        upk2(tb, tbx, tby); upk2(eb2, ebx, eby);
        em1 = fmax3abs(em1, ebx, eby);
        em1 = fmax3abs(em1, ez0, ez1);
        const unsigned addr_b = sel_b ? stg_addr<VoxT, true>(r, tbx, tby, tz1, full) : stg_addr<VoxT, false>(r, tbx, tby, tz1, full);
        const VoxT vb = lds_vox<VoxT>(addr_b);
        acc.add(va, cnt);
        acc.add(vb, cnt + 1);
    } else {
        em1 = fmax3abs(em1, ez0, ez0);
        acc.add(va, cnt);
    }
}

// The samples of cell (i, j, k) of the level-NB stream, corner (cxy, czz): its offsets in pairs.
template <int NB, typename VoxT, bool WANT_SQ, bool F32S>
DSC_D void stg_cell(const StgLane &r, bool full, StgAcc<VoxT, WANT_SQ, F32S> &acc, int i, int j, int k, f32x2 cxy, f32x2 czz, float &em, float &em1, int &cnt)
{
    const int s = i + j + k;
    const int m = (NB - s >= 3) ? 6 : (NB - s == 2 ? 5 : 1); // offsets kept in this cell
#pragma unroll
    for (int t = 0; t < 6; t += 2) {
        if (t < m) {
            const bool two = t + 1 < m;
            // z of the two samples in one packed add; x,y one packed add per sample
            f32x2 qzz = add2(czz, r.ozz[t >> 1]);
            f32x2 qa = add2(cxy, r.oxy[t]);
            f32x2 qb = two ? add2(cxy, r.oxy[t + 1]) : qa;
            if (NB >= 3) { // level-3 rows whose literals sum to 1 -/+ 1e-6 (every other level: eps = 0)
                const int ea = (NB == 3 || stg_need(s, t) <= 3) ? stg_eps3(i, j, k, t) : 0;
                const int eb = (two && (NB == 3 || stg_need(s, t + 1) <= 3)) ? stg_eps3(i, j, k, t + 1) : 0;
                if (ea > 0) qa = add2(qa, r.epxy);
                if (ea < 0) qa = sub2(qa, r.epxy);
                if (eb > 0) qb = add2(qb, r.epxy);
                if (eb < 0) qb = sub2(qb, r.epxy);
                if (ea != 0 || eb != 0) {
                    float ez0, ez1;
                    upk2(r.epzz, ez0, ez1);
                    qzz = add2(qzz, pk2(ea > 0 ? ez0 : (ea < 0 ? -ez0 : 0.f), eb > 0 ? ez1 : (eb < 0 ? -ez1 : 0.f)));
                }
            }
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, two, NB > 1 && stg_need(s, t) == NB, NB > 1 && two && stg_need(s, t + 1) == NB, cnt, em, em1);
            cnt += two ? 2 : 1;
        }
    }
}

// The level-NB instruction stream.  `full`: this lane's tet is of level NB (else NB-1, or the lane is idle with an all-zero
// record whose reads hit the first voxel of the slot and whose sums are never looked at).  Cells with i + j + k <= SMAX only.
template <int NB, typename VoxT, bool WANT_SQ, bool F32S, int SMAX = NB>
DSC_D void stg_cells(const StgLane &r, bool full, StgAcc<VoxT, WANT_SQ, F32S> &acc, float &em, float &em1)
{
    int cnt = 0; // samples so far (a constant at every point of the unrolled stream)
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const f32x2 cjxy = j ? fma2(pk2((float)j, (float)j), r.d1xy, r.q0xy) : r.q0xy;
        const f32x2 cjzz = j ? fma2(pk2((float)j, (float)j), r.d1zz, r.q0zz) : r.q0zz;
#pragma unroll
        for (int i = 0; i < NB - j; i++) {
            const f32x2 cixy = i ? fma2(pk2((float)i, (float)i), r.d2xy, cjxy) : cjxy;
            const f32x2 cizz = i ? fma2(pk2((float)i, (float)i), r.d2zz, cjzz) : cjzz;
#pragma unroll
            for (int k = 0; k < NB - j - i; k++) {
                if (i + j + k <= SMAX) {
                    const f32x2 cxy = k ? fma2(pk2((float)k, (float)k), r.d3xy, cixy) : cixy;
                    const f32x2 czz = k ? fma2(pk2((float)k, (float)k), r.d3zz, cizz) : cizz;
                    stg_cell<NB, VoxT, WANT_SQ, F32S>(r, full, acc, i, j, k, cxy, czz, em, em1, cnt);
                }
            }
        }
    }
}

template <int NB, typename VoxT, bool WANT_SQ, bool F32S> DSC_D void stg_block(const StgLane &r, bool full, StgAcc<VoxT, WANT_SQ, F32S> &acc)
{
    float em = 0.f, em1 = 0.f;
    stg_cells<NB, VoxT, WANT_SQ, F32S>(r, full, acc, em, em1);
    acc.emax = fmaxf(em, em1);
}

// The level-4 stream with its outer cells as loops: the unrolled stream is 17 KB of straight-line code, which together with the producer's
// code does not fit the SM's instruction cache.  The cell at P0 (6 samples) stays unrolled; the three cells with
// i + j + k = 1 (six offsets each), the six with i + j + k = 2 (five each) and the ten with i + j + k = 3 (one each) are loop bodies fed from
// constant tables.  The
// floating-point operations of a sample are those of stg_cells (a corner's skipped FMAs have a zero multiplier here: the same values).
struct StgCellRow { float j, i, k, e0; }; // cell indices, eps sign of its first offset (stg_eps3)
#define STG_ROW(I, J, K) {(float)(J), (float)(I), (float)(K), (float)stg_eps3(I, J, K, 0)}
struct StgCellRow6 { float j, i, k, e0, e1, e2, e3, e4; }; // i + j + k = 1: offsets 0..4 are level-3 rows as well
#define STG_ROW6(I, J, K) {(float)(J), (float)(I), (float)(K), (float)stg_eps3(I, J, K, 0), (float)stg_eps3(I, J, K, 1), (float)stg_eps3(I, J, K, 2), (float)stg_eps3(I, J, K, 3), (float)stg_eps3(I, J, K, 4)}
__constant__ StgCellRow6 c_stg4_s1[3] = {STG_ROW6(0, 0, 1), STG_ROW6(1, 0, 0), STG_ROW6(0, 1, 0)};
#undef STG_ROW6
__constant__ StgCellRow c_stg4_s2[6] = {STG_ROW(0, 0, 2), STG_ROW(1, 0, 1), STG_ROW(2, 0, 0), STG_ROW(0, 1, 1), STG_ROW(1, 1, 0), STG_ROW(0, 2, 0)};
__constant__ StgCellRow c_stg4_s3[10] = {STG_ROW(0, 0, 3), STG_ROW(1, 0, 2), STG_ROW(2, 0, 1), STG_ROW(3, 0, 0), STG_ROW(0, 1, 2),
                                         STG_ROW(1, 1, 1), STG_ROW(2, 1, 0), STG_ROW(0, 2, 1), STG_ROW(1, 2, 0), STG_ROW(0, 3, 0)};
#undef STG_ROW

template <typename VoxT, bool WANT_SQ, bool F32S> DSC_D void stg_block4_rolled(const StgLane &r, bool full, StgAcc<VoxT, WANT_SQ, F32S> &acc)
{
    float em = 0.f, em1 = 0.f;
    stg_cells<4, VoxT, WANT_SQ, F32S, 0>(r, full, acc, em, em1); // the cell at P0: six level-3 rows
    auto corner = [&](float j, float i, float k, f32x2 &cxy, f32x2 &czz) {
        const f32x2 j2 = pk2(j, j), i2 = pk2(i, i), k2 = pk2(k, k);
        cxy = fma2(k2, r.d3xy, fma2(i2, r.d2xy, fma2(j2, r.d1xy, r.q0xy)));
        czz = fma2(k2, r.d3zz, fma2(i2, r.d2zz, fma2(j2, r.d1zz, r.q0zz)));
    };
    // i + j + k = 1: six offsets; 0..4 are level-3 rows as well (eps), offset 5 exists at level 4 only
#pragma unroll 1
    for (int c = 0; c < 3; c++) {
        const StgCellRow6 row = c_stg4_s1[c];
        f32x2 cxy, czz;
        corner(row.j, row.i, row.k, cxy, czz);
        {
            const f32x2 qzz = fma2(pk2(row.e0, row.e1), r.epzz, add2(czz, r.ozz[0]));
            const f32x2 qa = fma2(pk2(row.e0, row.e0), r.epxy, add2(cxy, r.oxy[0])), qb = fma2(pk2(row.e1, row.e1), r.epxy, add2(cxy, r.oxy[1]));
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, true, false, false, 0, em, em1);
        }
        {
            const f32x2 qzz = fma2(pk2(row.e2, row.e3), r.epzz, add2(czz, r.ozz[1]));
            const f32x2 qa = fma2(pk2(row.e2, row.e2), r.epxy, add2(cxy, r.oxy[2])), qb = fma2(pk2(row.e3, row.e3), r.epxy, add2(cxy, r.oxy[3]));
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, true, false, false, 2, em, em1);
        }
        {
            const f32x2 qzz = fma2(pk2(row.e4, 0.f), r.epzz, add2(czz, r.ozz[2]));
            const f32x2 qa = fma2(pk2(row.e4, row.e4), r.epxy, add2(cxy, r.oxy[4])), qb = add2(cxy, r.oxy[5]);
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, true, false, true, 0, em, em1);
        }
    }
    // i + j + k = 2: offsets 0..4; offset 0 is a level-3 row as well (eps), offsets 1..4 exist at level 4 only
    constexpr int kUnrollS2 = STG_UNROLL_S2;
#pragma unroll kUnrollS2
    for (int c = 0; c < 6; c++) {
        const StgCellRow row = c_stg4_s2[c];
        f32x2 cxy, czz;
        corner(row.j, row.i, row.k, cxy, czz);
        {
            f32x2 qzz = add2(czz, r.ozz[0]);
            f32x2 qa = add2(cxy, r.oxy[0]);
            const f32x2 qb = add2(cxy, r.oxy[1]);
            qa = fma2(pk2(row.e0, row.e0), r.epxy, qa);       // +/- eps P0 (x, y), or nothing
            qzz = fma2(pk2(row.e0, 0.f), r.epzz, qzz);        // the same for z of the first sample
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, true, false, true, 0, em, em1);
        }
        {
            const f32x2 qzz = add2(czz, r.ozz[1]);
            const f32x2 qa = add2(cxy, r.oxy[2]), qb = add2(cxy, r.oxy[3]);
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qb, qzz, true, true, true, 2, em, em1);
        }
        {
            const f32x2 qzz = add2(czz, r.ozz[2]);
            const f32x2 qa = add2(cxy, r.oxy[4]);
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qa, qzz, false, true, false, 0, em, em1);
        }
    }
    // i + j + k = 3: offset 0 only, at level 4 only
#pragma unroll 1
    for (int c0 = 0; c0 < 10; c0 += 5) {
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const StgCellRow row = c_stg4_s3[c0 + u];
            f32x2 cxy, czz;
            corner(row.j, row.i, row.k, cxy, czz);
            const f32x2 qzz = add2(czz, r.ozz[0]);
            const f32x2 qa = add2(cxy, r.oxy[0]);
            stg_pair<VoxT, WANT_SQ, F32S>(r, full, acc, qa, qa, qzz, false, true, false, u & 3, em, em1);
        }
    }
    acc.emax = fmaxf(em, em1);
}

// |q_float - q_reference| per coordinate, for a tet whose origin-relative coordinates are below R and whose edge vectors
// from P0 are below diam (both in voxels):
//   q* = (P0 - O - 1/2) + j D1 + i D2 + k D3 + (oy E1 + ox E2 + oz E3) + eps P0 + tau,   tau = sum_{k>=1} eta_k (P_k - P0)
//   float: Q0, D_k, E_k = D_k/4 rounded once (2^-24 relative each); three FMAs for the cell corner (each rounds a number
//   below R + 1); the offset vector from three float operations on numbers below (9/4n) diam; one add for the sample and, at
//   level 3, one more for eps P0.  Sum: 2^-24 (6 R + 6 + 10 diam).  eta: the literals of levels 1, 2, 4 are exact binary
//   fractions; level 3 has |eta_k| <= 3.34e-7, so |tau| <= 1.01e-6 diam.  The reference's own FP64 rounding (4 products and 3
//   sums of numbers below the volume size) and the FP64 set-up of Q0, D_k add less than 1e-9.
DSC_D float stg_threshold(double R, double diam, int n)
{
    // evaluated in float, every step rounded towards a larger delta
    const float Rf = (float)R * 1.000001f, df = (float)diam * 1.000001f;
    const float delta = (5.9604644775390625e-08f * (6.0f * Rf + 6.0f + 14.0f * df) + (n == 3 ? 1.01e-6f * df : 0.0f) + 1.0e-9f) * 1.00001f;
    return (0.5f - delta) * 0.99999988079071044921875f;
}

DSC_D V3 load_pos_now(const double4 *pos4, unsigned node)
{
    const double4 *ptr = pos4 + node;
    V3 r;
    double pad;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%4];\n\tld.global.nc.v2.f64 {%2, %3}, [%4 + 16];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(pad) : "l"(ptr));
    return r;
}

// ---- shared-memory pipeline: mbarrier helpers ----
DSC_D void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
DSC_D void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
DSC_D void mbar_arrive_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
DSC_D void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// one bounded look: the warp is suspended by the hardware until the phase completes or a time limit passes
DSC_D bool mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

// Tets that leave the staged stream (level n >= 5, reaching outside the volume, a box that does not fit a slot, a sample inside the
// guard band) are appended to a list; tet_exact_list_kernel runs the reference's exact sequence on them afterwards.
DSC_D void exact_append(const TetArgs &a, bool flag, unsigned t, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(a.stg_exact_count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) a.stg_exact_list[base + __popc(m & ((1u << lane) - 1u))] = t;
}

// record of one tet in a slot: words [field][lane] (conflict-free), written by the producer lane, read by the consumer lane
enum { kRecQ0 = 0, kRecD1 = 3, kRecD2 = 6, kRecD3 = 9, kRecEp = 12, kRecThr = 15, kRecBase = 16, kRecMeta = 17, kRecW = 18, kRecV = 20, kRecWords = 22 };
constexpr int kStgRecBytes = kRecWords * 32 * 4;

// The kernel.  One CTA per SM: kStgProd producer warps and kStgCons consumer warps around a ring of shared-memory slots.
//   producer  lane <-> tet of a 32-tet tile (tiles fetched with one atomicAdd): label, ids, FP64 positions -> exact volume and level
//             (as the reference), voxel box of the tile; takes the next slot of the ring, writes the per-tet records of the FP32
//             lattice form, and has the TMA unit bring the box into the slot (one cp.async.bulk.tensor.3d, tensor map picked from a
//             family indexed by the box extents); the slot's "full" mbarrier completes when the bytes have landed.
//   consumer  takes the next filled slot, lane <-> tet again: the unrolled level-NB sample stream with every gather an LDS, the
//             guard-band test, per-label fixed-point sums in registers; releases the slot through its "empty" mbarrier.
// The phases of different tiles overlap: the dependent global loads of phase A, the TMA transfers and the sample streams of up to
// `stg_slots` tiles are in flight on one SM.
template <typename VoxT, bool WANT_SQ, bool PER_TET, bool F32S, int PMAX>
__global__ void __launch_bounds__((kStgProd + kStgCons) * 32, 1) tet_staged_kernel(const __grid_constant__ TetArgs a)
{
    extern __shared__ __align__(1024) unsigned char s_dyn[];
    __shared__ __align__(8) unsigned long long s_full[kStgMaxSlots]; // completes when the records are written and the box has landed
    __shared__ int4 s_hdr[kStgMaxSlots];         // tile base, NB (0: end of a producer's stream), row pitch, slice pitch
    __shared__ unsigned s_fillseq[kStgMaxSlots];  // sequence number + 1 of the item a producer has put into the slot
    __shared__ unsigned s_released[kStgMaxSlots]; // how many items of the slot the consumers are done with
    __shared__ __align__(8) unsigned long long s_empty[kStgMaxSlots]; // one phase per release: what a producer sleeps on (s_released is what it believes)
    __shared__ long long s_red[kStgCons][3 * kMaxPhase];
    __shared__ unsigned s_zero[4];
    __shared__ unsigned s_seq[2]; // next slot sequence number of the producers / of the consumers
    constexpr int ESZ = (int)sizeof(VoxT);
    constexpr int AL = 16 / ESZ; // elements per 16-byte unit
    const int X = a.vol.X, Y = a.vol.Y, Z = a.vol.Z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stg_slots;
    const unsigned s_rcp = 0xffffffffu / (unsigned)S + 1u; // q / S = umulhi(q, ceil(2^32 / S)) while q S < 2^32 (a CTA hands out fewer than 2^20 items, S <= 16)
    const unsigned box0 = smem_u32(s_dyn);
    const unsigned rec0 = box0 + (unsigned)S * (unsigned)a.stg_slot_bytes;
    if (threadIdx.x < 4) s_zero[threadIdx.x] = 0u;
    if (threadIdx.x < 2) s_seq[threadIdx.x] = 0u;
    if (threadIdx.x < S) {
        mbar_init(smem_u32(&s_full[threadIdx.x]), 1);
        mbar_init(smem_u32(&s_empty[threadIdx.x]), 1);
        s_fillseq[threadIdx.x] = 0u;
        s_released[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned n_work = a.tet_end - a.tet_begin;
    const unsigned n_tiles = (n_work + 31) / 32;

    if (warp < kStgProd) {
        // ======================================= producer =======================================
        unsigned long long nsamp = 0;
        // Next item of the ring and its slot, once the consumers are done with the slot's previous item.  Items are handed out in
        // sequence, but warps do not finish them in sequence, so a warp can be several rounds of the ring behind: the slot's history is
        // kept as counts (an mbarrier phase bit alone cannot tell round u from round u + 2).
        auto take_slot = [&](unsigned &slot, unsigned &q) {
            q = 0;
            if (lane == 0) q = atomicAdd(&s_seq[0], 1u);
            q = __shfl_sync(0xffffffffu, q, 0);
            const unsigned round = __umulhi(q, s_rcp);
            slot = q - round * (unsigned)S;
#ifdef STG_POLL_RELEASE
            while (*reinterpret_cast<volatile unsigned *>(&s_released[slot]) != round) __nanosleep(STG_PROD_SLEEP);
#else
            // the release counter says when; between two looks at it the warp sleeps on the slot's "empty" barrier, whose phase `round - 1`
            // completes with that release (a warp several rounds behind sees an aliased phase and simply looks again)
            while (*reinterpret_cast<volatile unsigned *>(&s_released[slot]) != round) mbar_try_wait(smem_u32(&s_empty[slot]), (round - 1u) & 1u);
#endif
            __threadfence_block();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the consumers' reads of the slot precede the TMA writes into it
        };
        // Three tiles are in flight per producer: the index of tile k+2 (atomicAdd), label and node ids of tile k+1, the positions of
        // tile k+1 (asked for at the top of iteration k, when its ids have arrived) -- so the dependent round trips of a tile
        // (index -> ids -> positions) overlap with the arithmetic of the tiles before it.
        // Tiles are claimed a.stg_chunk at a time (fewer when a launch has few tiles per warp: a chunk is also the grain of the tail): ptxas emits every returning atomic add in its warp-aggregated form, whose shuffle waits
        // for the result on the spot (one L2 round trip), so the claim is amortised over a chunk instead of being prefetched.
        // (inline PTX for the position loads: the compiler sinks plain loads down to their first use, an iteration later)
        const unsigned chunk = (unsigned)a.stg_chunk;
        unsigned chunk_base = 0, chunk_pos = chunk;
        auto fetch_tile = [&]() {
            if (chunk_pos == chunk) {
                unsigned x = 0;
                if (lane == 0) x = atomicAdd(a.stg_counter, chunk);
                chunk_base = __shfl_sync(0xffffffffu, x, 0);
                chunk_pos = 0;
            }
            return chunk_base + chunk_pos++;
        };
        auto load_ids = [&](unsigned tl, int &lab, uint4 &ids) {
            lab = 0; ids = make_uint4(0, 0, 0, 0);
            const unsigned tt = a.tet_begin + tl * 32 + lane;
            if (tl < n_tiles && tt < a.tet_end) { lab = __ldg(a.mesh.tet_label + tt); ids = __ldg(a.mesh.tet_nodes + tt); }
        };
        unsigned tile = fetch_tile();
        unsigned tile1 = fetch_tile();
        int label = 0, label1 = 0;
        uint4 nd, nd1;
        load_ids(tile, label, nd);
        load_ids(tile1, label1, nd1);
        V3 p0 = {0, 0, 0}, p1 = p0, p2 = p0, p3 = p0;
        if (tet_active(label, a.include_bound)) {
            p0 = load_pos(a.mesh.pos4, nd.x); p1 = load_pos(a.mesh.pos4, nd.y);
            p2 = load_pos(a.mesh.pos4, nd.z); p3 = load_pos(a.mesh.pos4, nd.w);
        }
        while (tile < n_tiles) {
            const unsigned tile2 = fetch_tile();
            // positions of the next tile (its ids are here by now)
            V3 n0 = {0, 0, 0}, n1 = n0, n2 = n0, n3 = n0;
            if (tet_active(label1, a.include_bound)) {
                n0 = load_pos_now(a.mesh.pos4, nd1.x); n1 = load_pos_now(a.mesh.pos4, nd1.y);
                n2 = load_pos_now(a.mesh.pos4, nd1.z); n3 = load_pos_now(a.mesh.pos4, nd1.w);
            }
            const unsigned base_t = a.tet_begin + tile * 32;
            const unsigned t = base_t + lane;
            const bool active = t < a.tet_end && tet_active(label, a.include_bound);
            bool cand = false;
            int L = 0;
            double v = 0;
            int blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0}, org[3] = {0, 0, 0};
            double rmax = 0, diam = 0;
            if (active) {
                v = tet_volume(p0, p1, p2, p3);
                L = tet_level(v);
                const double lox = fmin(fmin(p0.x, p1.x), fmin(p2.x, p3.x)), hix = fmax(fmax(p0.x, p1.x), fmax(p2.x, p3.x));
                const double loy = fmin(fmin(p0.y, p1.y), fmin(p2.y, p3.y)), hiy = fmax(fmax(p0.y, p1.y), fmax(p2.y, p3.y));
                const double loz = fmin(fmin(p0.z, p1.z), fmin(p2.z, p3.z)), hiz = fmax(fmax(p0.z, p1.z), fmax(p2.z, p3.z));
                // Inside the volume: no sample can leave [0, dim), so truncation equals floor and no bounds test is needed -- with a
                // margin that is plain; for tets with nodes on the faces of the volume see the argument at tet_integrate_w2_kernel
                // (minimum weight 1/(4n) - 5e-7 of every table row, checked when the table is built).
                auto axis_in = [](double lo, double hi, int dim) { return lo >= -1.0e-7 && hi <= (double)dim + 1.0e-7 && hi - lo >= 1.0e-4 * (double)dim; };
                const bool interior = (lox >= 1.5 && loy >= 1.5 && loz >= 1.5 && hix < (double)X - 1.5 && hiy < (double)Y - 1.5 && hiz < (double)Z - 1.5) ||
                                      (axis_in(lox, hix, X) && axis_in(loy, hiy, Y) && axis_in(loz, hiz, Z));
                const double ox = floor(lox), oy = floor(loy), oz = floor(loz);
                rmax = fmax(fmax(hix - ox, hiy - oy), hiz - oz);
                diam = fmax(fmax(hix - lox, hiy - loy), hiz - loz);
                cand = interior && L < 4 && rmax < 60.0;
                if (cand) {
                    // voxels the samples can touch: every weight is at least 1/(4n) - 5e-7, so a sample keeps (hi-lo)/(4n) away from the
                    // faces of the tet's bounding box; 0.9 of that and 1e-3 absolute cover eps * coordinate and the filter's own error
                    const double f = 0.9 / (double)(4 * (L + 1));
                    org[0] = (int)ox; org[1] = (int)oy; org[2] = (int)oz;
                    blo[0] = max((int)floor(lox + f * (hix - lox) - 1.0e-3), 0); bhi[0] = min((int)floor(hix - f * (hix - lox) + 1.0e-3), X - 1);
                    blo[1] = max((int)floor(loy + f * (hiy - loy) - 1.0e-3), 0); bhi[1] = min((int)floor(hiy - f * (hiy - loy) + 1.0e-3), Y - 1);
                    blo[2] = max((int)floor(loz + f * (hiz - loz) - 1.0e-3), 0); bhi[2] = min((int)floor(hiz - f * (hiz - loz) + 1.0e-3), Z - 1);
                }
                nsamp += (unsigned)((L + 1) * (L + 1) * (L + 1));
            } else if (PER_TET && t < a.tet_end) {
                if (a.tet_total) a.tet_total[t] = -1.0;
                if (a.tet_vol) a.tet_vol[t] = -1.0;
                if (a.tet_mean) a.tet_mean[t] = -1.0;
            }
            unsigned pending = __ballot_sync(0xffffffffu, cand);
            bool exact = active && !cand;
            while (pending) {
                unsigned set = pending;
                int x0 = 0, y0 = 0, z0 = 0, exu = 0, ey = 0, ez = 0;
                bool fits = false;
                for (;;) {
                    const bool in = (set >> lane) & 1u;
                    x0 = __reduce_min_sync(0xffffffffu, in ? blo[0] : INT_MAX);
                    y0 = __reduce_min_sync(0xffffffffu, in ? blo[1] : INT_MAX);
                    z0 = __reduce_min_sync(0xffffffffu, in ? blo[2] : INT_MAX);
                    const int x1 = __reduce_max_sync(0xffffffffu, in ? bhi[0] : INT_MIN);
                    const int y1 = __reduce_max_sync(0xffffffffu, in ? bhi[1] : INT_MIN);
                    const int z1 = __reduce_max_sync(0xffffffffu, in ? bhi[2] : INT_MIN);
                    x0 &= ~(AL - 1); // a TMA box starts on a 16-byte boundary
                    exu = (x1 - x0 + AL) / AL;
                    ey = y1 - y0 + 1;
                    ez = z1 - z0 + 1;
                    fits = exu <= kStgMaxUnits && ey <= kStgMaxRows && ez <= kStgMaxSlices && (unsigned)(exu * 16 * ey * ez) <= (unsigned)a.stg_slot_bytes;
                    const int cnt = __popc(set);
                    if (fits || cnt == 1) break;
                    set &= (1u << __fns(set, 0, cnt / 2 + 1)) - 1u; // keep the lower half of the set's lanes
                }
                pending &= ~set;
                const bool in = (set >> lane) & 1u;
                if (!fits) { exact = exact || in; continue; }
                // level of the instruction stream: the one that costs least (stream + the tets it leaves to the exact path)
                int NB;
                {
                    const unsigned m1 = __ballot_sync(0xffffffffu, in && L == 0), m2 = __ballot_sync(0xffffffffu, in && L == 1);
                    const unsigned m3 = __ballot_sync(0xffffffffu, in && L == 2), m4 = __ballot_sync(0xffffffffu, in && L == 3);
                    const int c1 = __popc(m1), c2 = __popc(m2), c3 = __popc(m3), c4 = __popc(m4);
                    const int k1 = 180, k4 = 260; // instruction estimates: exact path per tet (levels 1-3, level 4), streams below
                    const int cost1 = 40 + k1 * (c2 + c3) + k4 * c4, cost2 = 150 + k1 * c3 + k4 * c4, cost3 = 450 + k1 * c1 + k4 * c4, cost4 = 1000 + k1 * (c1 + c2);
                    NB = 1;
                    int best = c1 ? cost1 : INT_MAX;
                    if ((c2 || c1) && cost2 < best) { best = cost2; NB = 2; }
                    if ((c3 || c2) && cost3 < best) { best = cost3; NB = 3; }
                    if ((c4 || c3) && cost4 < best) { best = cost4; NB = 4; }
                }
                const bool ride = in && (L + 1 == NB || L + 2 == NB);
                exact = exact || (in && !ride);
                if (!__any_sync(0xffffffffu, ride)) continue;
                const TmapBlob *map = a.stg_maps + (((exu - 1) * kStgMaxRows + (ey - 1)) * kStgMaxSlices + (ez - 1));
#ifndef STG_NO_L2_PREFETCH
                // the box on its way into L2 while this warp waits for a slot: the load into the slot then finds most of it there
                if (lane == 0)
                    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x0), "r"(y0), "r"(z0) : "memory");
#endif
                unsigned slot, q;
                take_slot(slot, q);
                const unsigned slot_a = box0 + slot * (unsigned)a.stg_slot_bytes;
                const unsigned bx = (unsigned)(exu * 16), bxy = bx * (unsigned)ey;
                // the box first: its bytes travel while the records are written (the consumer needs the records before the voxels)
                if (lane == 0) {
                    const unsigned bar = smem_u32(&s_full[slot]);
                    mbar_arrive_tx(bar, bxy * (unsigned)ez);
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(slot_a), "l"(map),
                                 "r"(x0), "r"(y0), "r"(z0), "r"(bar)
                                 : "memory");
                }
                {
                    float *rec = reinterpret_cast<float *>(s_dyn + (size_t)S * a.stg_slot_bytes + (size_t)slot * kStgRecBytes) + lane;
                    const int n = L + 1;
                    const double inv = n == 1 ? 1.0 : (n == 2 ? 0.5 : (n == 3 ? 1.0 / 3.0 : 0.25)); // = 1.0 / n, correctly rounded
                    rec[32 * (kRecQ0 + 0)] = ride ? (float)((p0.x - (double)org[0]) - 0.5) : 0.f;
                    rec[32 * (kRecQ0 + 1)] = ride ? (float)((p0.y - (double)org[1]) - 0.5) : 0.f;
                    rec[32 * (kRecQ0 + 2)] = ride ? (float)((p0.z - (double)org[2]) - 0.5) : 0.f;
                    rec[32 * (kRecD1 + 0)] = ride ? (float)((p1.x - p0.x) * inv) : 0.f; rec[32 * (kRecD1 + 1)] = ride ? (float)((p1.y - p0.y) * inv) : 0.f;
                    rec[32 * (kRecD1 + 2)] = ride ? (float)((p1.z - p0.z) * inv) : 0.f;
                    rec[32 * (kRecD2 + 0)] = ride ? (float)((p2.x - p0.x) * inv) : 0.f; rec[32 * (kRecD2 + 1)] = ride ? (float)((p2.y - p0.y) * inv) : 0.f;
                    rec[32 * (kRecD2 + 2)] = ride ? (float)((p2.z - p0.z) * inv) : 0.f;
                    rec[32 * (kRecD3 + 0)] = ride ? (float)((p3.x - p0.x) * inv) : 0.f; rec[32 * (kRecD3 + 1)] = ride ? (float)((p3.y - p0.y) * inv) : 0.f;
                    rec[32 * (kRecD3 + 2)] = ride ? (float)((p3.z - p0.z) * inv) : 0.f;
                    const bool l3 = ride && n == 3;
                    rec[32 * (kRecEp + 0)] = l3 ? (float)(1.0e-6 * p0.x) : 0.f;
                    rec[32 * (kRecEp + 1)] = l3 ? (float)(1.0e-6 * p0.y) : 0.f;
                    rec[32 * (kRecEp + 2)] = l3 ? (float)(1.0e-6 * p0.z) : 0.f;
                    rec[32 * kRecThr] = ride ? stg_threshold(rmax + 1.0, diam, n) : 1.0e30f;
                    // address of voxel O + m: base + (bits(tx) << SH) + bits(ty) bx + bits(tz) bxy, the bits being 0x4B400000 + m
                    const unsigned fold = 0x4B400000u * ((unsigned)ESZ + bx + bxy);
                    const unsigned rel = ride ? (unsigned)((org[0] - x0) * ESZ) + (unsigned)(org[1] - y0) * bx + (unsigned)(org[2] - z0) * bxy : 0u;
                    unsigned *recu = reinterpret_cast<unsigned *>(rec);
                    recu[32 * kRecBase] = slot_a + rel - fold;
                    recu[32 * kRecMeta] = (ride ? 1u : 0u) | ((ride && L + 1 == NB) ? 2u : 0u) | ((unsigned)(label & 0xff) << 8) | ((unsigned)L << 16);
                    const double w = v / (double)(n * n * n); // the same expression as in tet_exact_list_kernel: a tet contributes the same integer whichever path it takes
                    recu[32 * kRecW] = (unsigned)__double2loint(w); recu[32 * (kRecW + 1)] = (unsigned)__double2hiint(w);
                    recu[32 * kRecV] = (unsigned)__double2loint(v); recu[32 * (kRecV + 1)] = (unsigned)__double2hiint(v);
                    if (lane == 0) s_hdr[slot] = make_int4((int)base_t, NB, (int)bx, (int)bxy);
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    *reinterpret_cast<volatile unsigned *>(&s_fillseq[slot]) = q + 1u;
                }
            }
            exact_append(a, exact, t, lane);
            // shift the pipeline: ids of the tile after next
            tile = tile1; label = label1; nd = nd1;
            p0 = n0; p1 = n1; p2 = n2; p3 = n3;
            tile1 = tile2;
            load_ids(tile1, label1, nd1);
        }
        // end of this producer's stream: one END item per consumer it is responsible for
        for (int e = warp; e < kStgCons; e += kStgProd) {
            unsigned slot, q;
            take_slot(slot, q);
            if (lane == 0) {
                s_hdr[slot] = make_int4(0, 0, 0, 0);
                __threadfence_block();
                *reinterpret_cast<volatile unsigned *>(&s_fillseq[slot]) = q + 1u;
                mbar_arrive(smem_u32(&s_full[slot]));
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, m);
        if (lane == 0 && nsamp) atomicAdd(a.sample_count, nsamp);
    } else {
        // ======================================= consumer =======================================
        long long acc_tot[PMAX], acc_sq[PMAX], acc_vol[PMAX]; // labels 1..PMAX (the host picks PMAX >= nb_phase: 4 or 8)
#pragma unroll
        for (int k = 0; k < PMAX; k++) acc_tot[k] = acc_sq[k] = acc_vol[k] = 0;
        for (;;) {
            unsigned q = 0;
            if (lane == 0) q = atomicAdd(&s_seq[1], 1u);
            q = __shfl_sync(0xffffffffu, q, 0);
            const unsigned round = __umulhi(q, s_rcp), slot = q - round * (unsigned)S;
            // first the producer of THIS item must have started filling the slot (then the mbarrier is in the phase of this round) ...
            while (*reinterpret_cast<volatile unsigned *>(&s_fillseq[slot]) != q + 1u) __nanosleep(40);
            __threadfence_block(); // the records and the header were written before the sequence number
            // ... then its phase completes when the voxels have landed: waited for after the per-lane set-up below
            const int4 hdr = s_hdr[slot];
            const int NB = hdr.y;
            if (NB == 0) { // a producer has finished
                __syncwarp();
                if (lane == 0) {
                    *reinterpret_cast<volatile unsigned *>(&s_released[slot]) = round + 1u;
                    mbar_arrive(smem_u32(&s_empty[slot]));
                }
                break;
            }
            const float *rec = reinterpret_cast<const float *>(s_dyn + (size_t)S * a.stg_slot_bytes + (size_t)slot * kStgRecBytes) + lane;
            const unsigned *recu = reinterpret_cast<const unsigned *>(rec);
            StgLane r;
            r.bx = (unsigned)hdr.z;
            r.bxy = (unsigned)hdr.w;
            r.zero_addr = smem_u32(&s_zero[0]);
            const unsigned meta = recu[32 * kRecMeta];
            const bool ride = meta & 1u, full = (meta & 2u) != 0;
            const int label = (int)((meta >> 8) & 0xffu), L = (int)(meta >> 16);
            {
                const float q0x = rec[32 * (kRecQ0 + 0)], q0y = rec[32 * (kRecQ0 + 1)], q0z = rec[32 * (kRecQ0 + 2)];
                const float d1x = rec[32 * (kRecD1 + 0)], d1y = rec[32 * (kRecD1 + 1)], d1z = rec[32 * (kRecD1 + 2)];
                const float d2x = rec[32 * (kRecD2 + 0)], d2y = rec[32 * (kRecD2 + 1)], d2z = rec[32 * (kRecD2 + 2)];
                const float d3x = rec[32 * (kRecD3 + 0)], d3y = rec[32 * (kRecD3 + 1)], d3z = rec[32 * (kRecD3 + 2)];
                r.q0xy = pk2(q0x, q0y); r.q0zz = pk2(q0z, q0z);
                r.d1xy = pk2(d1x, d1y); r.d1zz = pk2(d1z, d1z);
                r.d2xy = pk2(d2x, d2y); r.d2zz = pk2(d2z, d2z);
                r.d3xy = pk2(d3x, d3y); r.d3zz = pk2(d3z, d3z);
                // centroid offsets oy E1 + ox E2 + oz E3, E_k = D_k / 4: quarters are exact, so (o/4) D_k in the FMA is the same product
                const f32x2 Z2 = pk2(0.f, 0.f);
#pragma unroll
                for (int o = 0; o < 6; o++) {
                    const float fx = 0.25f * (float)stg_off(o, 0), fy = 0.25f * (float)stg_off(o, 1), fz = 0.25f * (float)stg_off(o, 2);
                    r.oxy[o] = fma2(pk2(fz, fz), r.d3xy, fma2(pk2(fx, fx), r.d2xy, fma2(pk2(fy, fy), r.d1xy, Z2)));
                }
#pragma unroll
                for (int o = 0; o < 6; o += 2) { // z of offsets (o, o + 1) in the two halves
                    const f32x2 cx = pk2(0.25f * (float)stg_off(o, 0), 0.25f * (float)stg_off(o + 1, 0)), cy = pk2(0.25f * (float)stg_off(o, 1), 0.25f * (float)stg_off(o + 1, 1));
                    const f32x2 cz = pk2(0.25f * (float)stg_off(o, 2), 0.25f * (float)stg_off(o + 1, 2));
                    r.ozz[o >> 1] = fma2(cz, r.d3zz, fma2(cx, r.d2zz, fma2(cy, r.d1zz, Z2)));
                }
                const float epz = rec[32 * (kRecEp + 2)];
                r.epxy = pk2(rec[32 * (kRecEp + 0)], rec[32 * (kRecEp + 1)]); r.epzz = pk2(epz, epz);
                r.thr = rec[32 * kRecThr];
                r.base = recu[32 * kRecBase];
            }
            const double w = __hiloint2double((int)recu[32 * (kRecW + 1)], (int)recu[32 * kRecW]);
            const double v = __hiloint2double((int)recu[32 * (kRecV + 1)], (int)recu[32 * kRecV]);
            StgAcc<VoxT, WANT_SQ, F32S> acc;
            mbar_wait(smem_u32(&s_full[slot]), round & 1u);
            switch (NB) {
            case 1: stg_block<1, VoxT, WANT_SQ, F32S>(r, full, acc); break;
            case 2: stg_block<2, VoxT, WANT_SQ, F32S>(r, full, acc); break;
            case 3: stg_block<3, VoxT, WANT_SQ, F32S>(r, full, acc); break;
#ifndef STG_UNROLLED
            default: stg_block4_rolled<VoxT, WANT_SQ, F32S>(r, full, acc); break;
#else
            default: stg_block<4, VoxT, WANT_SQ, F32S>(r, full, acc); break; // (A/B: the fully unrolled level-4 stream)
#endif
            }
            __syncwarp();
            if (lane == 0) { // every lane has read its voxels and its record
                __threadfence_block();
                *reinterpret_cast<volatile unsigned *>(&s_released[slot]) = round + 1u;
                mbar_arrive(smem_u32(&s_empty[slot])); // (release semantics: the counter is visible to whoever sees the phase complete)
            }
            const bool flagged = ride && !(acc.emax < r.thr);
            const unsigned t = (unsigned)hdr.x + (unsigned)lane;
            exact_append(a, flagged, t, lane);
            if (ride && !flagged) {
                double t_sum, t_sq;
                acc.result(t_sum, t_sq);
                const int n3 = (L + 1) * (L + 1) * (L + 1);
                const double tot = PER_TET ? t_sum * v / (double)n3 : t_sum * w; // total * v / a.size() (image3d.cpp:371)
                if (PER_TET) {
                    if (a.tet_total) a.tet_total[t] = tot;
                    if (a.tet_vol) a.tet_vol[t] = v;
                    if (a.tet_mean) a.tet_mean[t] = tot / v;
                }
                const long long ftot = __double2ll_rn(tot * a.stg_scale[0]);
                const long long fvol = __double2ll_rn(v * a.stg_scale[2]);
                long long fsq = 0;
                if (WANT_SQ) fsq = __double2ll_rn(t_sq * w * a.stg_scale[1]);
#pragma unroll
                for (int k = 0; k < PMAX; k++)
                    if (label == k + 1 && k < a.nb_phase) {
                        acc_tot[k] += ftot;
                        acc_vol[k] += fvol;
                        if (WANT_SQ) acc_sq[k] += fsq;
                    }
            }
        }
        // per-warp rows (integers: any order gives the same sums)
#pragma unroll
        if (lane < 3 * kMaxPhase) s_red[warp - kStgProd][lane] = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < PMAX; k++) {
            long long x = acc_tot[k], y = acc_sq[k], z = acc_vol[k];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                x += __shfl_xor_sync(0xffffffffu, x, m);
                if (WANT_SQ) y += __shfl_xor_sync(0xffffffffu, y, m);
                z += __shfl_xor_sync(0xffffffffu, z, m);
            }
            if (lane == 0) {
                s_red[warp - kStgProd][k] = x;
                s_red[warp - kStgProd][kMaxPhase + k] = y;
                s_red[warp - kStgProd][2 * kMaxPhase + k] = z;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * kMaxPhase) {
        long long rsum = 0;
        for (int wd = 0; wd < kStgCons; wd++) rsum += s_red[wd][threadIdx.x];
        if (rsum) atomicAdd(a.stg_acc + threadIdx.x, (unsigned long long)rsum); // integers: the order of the CTAs does not matter
    }
}

// The tets the staged stream left out (exact_append), one warp per tet: the reference's exact FP64 sequence for every sample
// (get_coord, DEMO/tet_dis_coord.hpp:27; truncation image3d.cpp:366; bounds image3d.cpp:480-491), two table rows per lane in
// flight.  The list entry and the ids of the following tets are fetched ahead.
template <typename VoxT, bool WANT_SQ, bool PER_TET, bool F32S>
__global__ void __launch_bounds__(256, 4) tet_exact_list_kernel(const __grid_constant__ TetArgs a)
{
    __shared__ unsigned long long s_acc[3 * kMaxPhase];
    if (threadIdx.x < 3 * kMaxPhase) s_acc[threadIdx.x] = 0ull;
    __syncthreads();
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    const int X = a.vol.X, Y = a.vol.Y, Z = a.vol.Z;
    const int lane = threadIdx.x & 31;
    const unsigned n = *a.stg_exact_count;
    const unsigned done = a.stg_exact_done ? *a.stg_exact_done : 0u; // entries an earlier launch of this pass has worked off
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    constexpr unsigned NONE = 0xffffffffu;
    auto entry = [&](unsigned i) { return i < n ? __ldg(a.stg_exact_list + i) : NONE; };
    auto ids = [&](unsigned t, int &lab, uint4 &nd) {
        lab = 0; nd = make_uint4(0, 0, 0, 0);
        if (t != NONE) { lab = __ldg(a.mesh.tet_label + t); nd = __ldg(a.mesh.tet_nodes + t); }
    };
    unsigned i = done + wid;
    unsigned t = entry(i), t1 = entry(i + nw), t2 = NONE;
    int label, label1;
    uint4 nd, nd1;
    ids(t, label, nd);
    ids(t1, label1, nd1);
    while (t != NONE) {
        t2 = entry(i + 2 * nw);
        const V3 p0 = load_pos(a.mesh.pos4, nd.x), p1 = load_pos(a.mesh.pos4, nd.y), p2 = load_pos(a.mesh.pos4, nd.z), p3 = load_pos(a.mesh.pos4, nd.w);
        const double v = tet_volume(p0, p1, p2, p3);
        const int L = tet_level(v);
        const int rows = (L + 1) * (L + 1) * (L + 1);
        const int off = (L * (L + 1) / 2) * (L * (L + 1) / 2);
        const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(a.tet_bary) + 2 * (size_t)off;
        StgAcc<VoxT, WANT_SQ, F32S> acc;
        constexpr int U = 2; // table rows per lane in flight
        for (int r0 = 0; r0 < rows; r0 += 32 * U) {
            double2 b01[U], b23[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int r = min(r0 + 32 * u + lane, rows - 1);
                b01[u] = __ldg(tb + 2 * r);
                b23[u] = __ldg(tb + 2 * r + 1);
            }
            VoxT val[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ix = (int)(((p0.x * b01[u].x + p1.x * b01[u].y) + p2.x * b23[u].x) + p3.x * b23[u].y);
                const int iy = (int)(((p0.y * b01[u].x + p1.y * b01[u].y) + p2.y * b23[u].x) + p3.y * b23[u].y);
                const int iz = (int)(((p0.z * b01[u].x + p1.z * b01[u].y) + p2.z * b23[u].x) + p3.z * b23[u].y);
                const bool ok = r0 + 32 * u + lane < rows && (unsigned)ix < (unsigned)X && (unsigned)iy < (unsigned)Y && (unsigned)iz < (unsigned)Z;
                val[u] = ok ? __ldg(vol + ((size_t)iz * Y + iy) * X + ix) : (VoxT)0; // outside the volume: BOUND_INTENSITY = 0
            }
#pragma unroll
            for (int u = 0; u < U; u++) acc.add(val[u], u);
        }
        double s, q;
        if constexpr (VoxTraits<VoxT>::integral) {
            // integers first, one division: the same number the staged stream forms for a tet (its lane adds all samples of its tet)
            const unsigned is = __reduce_add_sync(0xffffffffu, acc.isum[0] + acc.isum[1]); // <= 729 * 65535
            unsigned long long iq = acc.isq;
            if (WANT_SQ) {
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) iq += __shfl_xor_sync(0xffffffffu, iq, m);
            }
            const double dn = VoxTraits<VoxT>::denom();
            s = (double)is / dn;
            q = WANT_SQ ? (double)iq / (dn * dn) : 0.0;
        } else {
            acc.result(s, q);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                s += shfl_xor_d(s, m);
                if (WANT_SQ) q += shfl_xor_d(q, m);
            }
        }
        if (lane == 0) {
            const double w = v / (double)rows;
            const double tot = PER_TET ? s * v / (double)rows : s * w;
            if (PER_TET) {
                if (a.tet_total) a.tet_total[t] = tot;
                if (a.tet_vol) a.tet_vol[t] = v;
                if (a.tet_mean) a.tet_mean[t] = tot / v;
            }
            if (label >= 1 && label <= a.nb_phase) {
                atomicAdd(&s_acc[label - 1], (unsigned long long)__double2ll_rn(tot * a.stg_scale[0]));
                if (WANT_SQ) atomicAdd(&s_acc[kMaxPhase + label - 1], (unsigned long long)__double2ll_rn(q * w * a.stg_scale[1]));
                atomicAdd(&s_acc[2 * kMaxPhase + label - 1], (unsigned long long)__double2ll_rn(v * a.stg_scale[2]));
            }
        }
        i += nw;
        t = t1; label = label1; nd = nd1;
        t1 = t2;
        ids(t1, label1, nd1);
    }
    __syncthreads();
    if (threadIdx.x < 3 * kMaxPhase && s_acc[threadIdx.x]) atomicAdd(a.stg_acc + threadIdx.x, s_acc[threadIdx.x]);
}

// The fixed-point sums (one row of 3 * kMaxPhase integers the CTAs of the staged kernels have added their sums to) -> doubles:
// out[0..P) total, [P..2P) sumsq, [2P..3P) volume, mean = total/volume.  Leaves the row, the tile counter and the list length zero
// for the next pass.  A sum beyond 2^62 (a mesh far larger than the volume the scale was chosen for) raises bit 1 of *flags.
__global__ void __launch_bounds__(32) tet_finalize_fixed_kernel(unsigned long long *acc, int nb_phase, double s_tot, double s_sq, double s_vol, double *sums, double *mean,
                                                                unsigned *flags, unsigned *stg_counters)
{
    __shared__ double tot[kMaxPhase], vol[kMaxPhase];
    const int col = threadIdx.x;
    if (col < 4 && stg_counters) stg_counters[col] = 0u;
    if (col < 3 * kMaxPhase) {
        const long long r = (long long)acc[col];
        acc[col] = 0ull;
        const int which = col / kMaxPhase, ph = col % kMaxPhase;
        const double scale = which == 0 ? s_tot : (which == 1 ? s_sq : s_vol);
        const double val = (double)r / scale; // scale is a power of two: one rounding, of the integer
        if (fabs((double)r) >= 4.0e18 && flags) atomicOr(flags, 2u);
        if (ph < nb_phase) sums[which * nb_phase + ph] = val;
        if (which == 0) tot[ph] = val;
        if (which == 2) vol[ph] = val;
    }
    __syncthreads();
    if (mean && threadIdx.x < nb_phase) mean[threadIdx.x] = tot[threadIdx.x] / vol[threadIdx.x];
}

} // namespace dsc
