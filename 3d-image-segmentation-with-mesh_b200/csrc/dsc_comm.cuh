// dsc_comm.cuh -- the two exchange steps of a multi-GPU image-force evaluation as kernels over NVLink peer memory.
//
// every GPU holds the whole volume and snapshot and a partition of the tets and interface faces; the only
// cross-GPU data are 3*nb_phase per-phase sums (needed by the force pass) and the force rows of the nodes that have
// interface faces (14 % of the nodes at C4: 1.15 MB).  Both are far below the size where NCCL's bandwidth matters; what a
// step pays is NCCL's launch and synchronisation latency, twice, plus the pack / unpack kernels around the second call --
// about 0.1 ms of a 0.19 ms step at 8 GPUs.  Here the exchanges are PUSHES: a rank stores its contribution into a slot of
// every peer's window (posted writes over NVLink: no round trip), then stores its flag there; a rank only ever polls and reads
// its OWN memory.  (The first form pulled: every thread read eight peers' rows with dependent remote loads and every block
// polled remote flags -- 0.09 ms for the force exchange at 8 GPUs, slower than NCCL.)
//
//   comm_phase_allreduce_kernel  adds this rank's per-CTA partial rows of the tet pass (what tet_finalize_kernel does),
//                                pushes the 3*nb_phase sums and the flag to every peer, waits for all flags in its own
//                                window, adds the slots in rank order (same order on every rank: identical results) and
//                                writes sums and means.
//   comm_force_fused_kernel      the whole force exchange in one launch: forms the vectors of this rank's faces from the
//                                mean-independent geometry and scatter-adds them into compact local rows, pushes the range of
//                                rows it touched into its slot of every window, then its flag; waits for all flags (local
//                                polling) and adds, row by row, the slots of the ranks that touched the row, in rank order.
//
// Windows are cudaMalloc'ed per rank and mapped into the peers with CUDA IPC (dscgpu_comm_create / dscgpu_comm_connect).
// Each window has two halves used by alternating evaluations (epoch parity): a rank can be one exchange ahead of a slow peer,
// never two, because it cannot pass exchange e+1 before that peer has pushed its flag e+1, which it does after its reads of
// exchange e.  Flags carry the epoch (monotonic), so nothing is ever reset.  The epoch lives in device memory and is advanced
// by the kernel itself: the step can be captured in a CUDA graph and replayed.
// A wait that sees no progress for ~2 s sets an error word instead of hanging the device.
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

constexpr int kCommMaxWorld = 8;
// window layout (bytes): flags of the sums exchange [2 halves][8 ranks] x 8 B at 0; of the force exchange at 128; epoch 256;
// error 264; push-kernel block counter 272; sums slots [2][8][24] doubles at 512 (3072 B); rows slots [2][world][3 cap_rows] at 4096
constexpr size_t kCommFlagSums = 0, kCommFlagForce = 128, kCommEpoch = 256, kCommError = 264, kCommDone = 272, kCommSums = 512, kCommRows = 4096;
// fused force exchange: row ranges pushed by the ranks [2 halves][8 ranks] x (lo, hi) at 3584; local scratch [2 halves] x (barrier 1, barrier 2, lo, hi) at 3712
constexpr size_t kCommRanges = 3584, kCommLocal = 3712;
// sharded snapshot commit: data flags [8] at 320, acknowledgements [8] at 384, its own epoch at 448, push-kernel block counter at 456
constexpr size_t kCommFlagSnap = 320, kCommAckSnap = 384, kCommSnapEpoch = 448, kCommSnapDone = 456;

struct CommView {
    unsigned char *win[kCommMaxWorld]; // window of every rank as mapped here; win[rank] is this rank's own
    int rank, world;
    unsigned cap_rows;
    size_t snap_off; // byte offset of the snapshot region of a window (after the row slots)
};

DSC_HD size_t comm_snapshot_offset(int world, unsigned cap_rows) { return (kCommRows + 2 * (size_t)world * 3 * cap_rows * sizeof(double) + 255) & ~(size_t)255; }
DSC_HD size_t comm_window_bytes(int world, unsigned cap_rows, size_t snap_bytes = 0) { return comm_snapshot_offset(world, cap_rows) + snap_bytes; }
DSC_HD size_t comm_rows_offset(int half, int slot, int world, unsigned cap_rows) { return kCommRows + ((size_t)half * world + slot) * 3 * cap_rows * sizeof(double); }

DSC_D unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DSC_D void st_release_sys(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// waits until the (local) flag at p reaches epoch e; false (and the error word set) after ~2 s without success
DSC_D bool comm_wait(const unsigned long long *p, unsigned long long e, int *err)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < e) {
        if (clock64() - t0 > 4000000000ll) {
            *err = 1;
            return false;
        }
    }
    return true;
}

// FIXED: the partial rows are the 64-bit fixed-point sums of the staged tet kernel (dsc_tet_staged.cuh): integers are exchanged and
// added, so the result does not depend on the number of ranks nor on how the tets were cut; scale[which] converts at the end.
template <bool FIXED>
__global__ void __launch_bounds__(32 * 3 * kMaxPhase) comm_phase_allreduce_kernel(CommView c, const void *__restrict__ partials_v, int n_rows, int nb_phase, double s_tot,
                                                                                  double s_sq, double s_vol, double *sums, double *mean, unsigned *stg_counters)
{
    if (FIXED && threadIdx.x < 4 && stg_counters) stg_counters[threadIdx.x] = 0u; // staged kernel: tile counter, exact-list length, worked-off part
    const int col = threadIdx.x >> 5, lane = threadIdx.x & 31; // 24 warps, one per column of the partial rows
    unsigned char *mine = c.win[c.rank];
    __shared__ unsigned long long s_epoch;
    __shared__ double tot[kMaxPhase], vol[kMaxPhase];
    if (threadIdx.x == 0) {
        unsigned long long *ep = reinterpret_cast<unsigned long long *>(mine + kCommEpoch);
        s_epoch = *ep + 1; // one epoch per evaluation; the force exchange of the same evaluation reads it back
        *ep = s_epoch;
    }
    unsigned long long bits; // what is pushed: the double, or the integer, of this column
    if (FIXED) { // one row of integers, the CTAs of the staged kernels have added theirs already; left zero for the next pass
        unsigned long long *acc = static_cast<unsigned long long *>(const_cast<void *>(partials_v));
        bits = acc[col];
        __syncwarp();
        if (lane == 0) acc[col] = 0ull;
    } else {
        const double *partials = static_cast<const double *>(partials_v);
        double r = 0;
        for (int i = lane; i < n_rows; i += 32) r += partials[(size_t)i * 3 * kMaxPhase + col];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) r += shfl_xor_d(r, m);
        bits = (unsigned long long)__double_as_longlong(r);
    }
    __syncthreads();
    const unsigned long long e = s_epoch;
    const int half = (int)(e & 1);
    // push: lane k of every column warp stores the column's sum into this rank's slot of rank k's window
    if (lane < c.world) reinterpret_cast<unsigned long long *>(c.win[lane] + kCommSums)[((size_t)half * kCommMaxWorld + c.rank) * 3 * kMaxPhase + col] = bits;
    __threadfence_system();
    __syncthreads();
    __shared__ int s_failed;
    if (threadIdx.x == 0) s_failed = 0;
    __syncthreads();
    if (threadIdx.x < c.world) {
        st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommFlagSums) + half * kCommMaxWorld + c.rank, e);
        if (!comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommFlagSums) + half * kCommMaxWorld + threadIdx.x, e, reinterpret_cast<int *>(mine + kCommError)))
            s_failed = 1;
    }
    __syncthreads();
    if (s_failed) { // a peer never arrived: no partial result may pass for a sum (dscgpu_comm_error reports the timeout)
        const int which = col / kMaxPhase, ph = col % kMaxPhase;
        if (lane == 0 && ph < nb_phase) sums[which * nb_phase + ph] = __longlong_as_double(0x7ff8000000000000ll);
        if (mean && threadIdx.x < nb_phase) mean[threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    if (lane == 0) {
        const int which = col / kMaxPhase, ph = col % kMaxPhase;
        const unsigned long long *slot = reinterpret_cast<const unsigned long long *>(mine + kCommSums) + (size_t)half * kCommMaxWorld * 3 * kMaxPhase + col;
        double t = 0;
        if (FIXED) {
            long long ti = 0;
            for (int k = 0; k < c.world; k++) ti += (long long)__ldcv(slot + (size_t)k * 3 * kMaxPhase);
            t = (double)ti / (which == 0 ? s_tot : (which == 1 ? s_sq : s_vol)); // power of two: one rounding, of the integer
        } else {
            for (int k = 0; k < c.world; k++) t += __longlong_as_double((long long)__ldcv(slot + (size_t)k * 3 * kMaxPhase));
        }
        if (ph < nb_phase) sums[which * nb_phase + ph] = t;
        if (which == 0) tot[ph] = t;
        if (which == 2) vol[ph] = t;
    }
    __syncthreads();
    if (mean && threadIdx.x < nb_phase) mean[threadIdx.x] = tot[threadIdx.x] / vol[threadIdx.x];
}

// ---- the force exchange as ONE kernel -------------------------------------------------------------------------------------
// comm_force_fused_kernel: (1) forms the force vectors of this rank's interface faces from the mean-independent geometry
// (force_geom_kernel: normals, areas, trilinear samples) and the phase means, scatter-adds them into compact local rows and
// notes the range of rows it touched; zeroes the full force array; (2) after a grid-wide barrier pushes ONLY that range into its
// slot of every peer's window (a rank's faces touch about 1/world of the rows: the dense form pushed every row to every peer),
// together with the range, then -- once its last block has pushed -- its flag; (3) waits for the peers' flags in its own window and
// adds, for every row, the slots of the ranks whose range holds it, in rank order, into the full force array.
// The grid must be co-resident (it spins on its own barrier counters): the host sizes it from the occupancy query.
struct ForceFusedArgs {
    CommView c;
    MeshView mesh;
    const double *fgeo;
    const int *flab;
    const double *fint;
    const signed char *fset;
    int nb_phase;
    const double *mean;
    const unsigned *apos;         // node -> row (exclusive scan of "has interface faces")
    double *rows;                 // local rows, 3 per active node, zero on entry and on exit
    unsigned begin, end;          // faces of this rank
    const unsigned *active_nodes; // row -> node
    const unsigned *n_active_ptr;
    double *forces;               // full array, 3 * n_nodes
    unsigned n_nodes;
};

DSC_D void comm_grid_barrier(unsigned *counter)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned *>(counter) < gridDim.x) { }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) comm_force_fused_kernel(const ForceFusedArgs a)
{
    const CommView &c = a.c;
    unsigned char *mine = c.win[c.rank];
    const unsigned long long e = *reinterpret_cast<const unsigned long long *>(mine + kCommEpoch);
    const int half = (int)(e & 1);
    unsigned *loc = reinterpret_cast<unsigned *>(mine + kCommLocal) + 4 * half; // barrier 1, barrier 2, lo, hi
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // (1) zero the full array, scatter this rank's faces into the local rows
    for (size_t i = gtid; i < 3 * (size_t)a.n_nodes; i += gsz) a.forces[i] = 0.0;
    unsigned lo = 0xffffffffu, hi = 0u;
    for (unsigned f = a.begin + gtid; f < a.end; f += gsz) {
        const int gs = a.fset[f];
        if (gs < 0) continue;
        const double *g = a.fgeo + 4 * (size_t)f;
        const V3 norm = {g[0], g[1], g[2]};
        const double area = g[3];
        const double c0 = phase_intensity(a.flab[2 * (size_t)f], a.nb_phase, a.mean), c1 = phase_intensity(a.flab[2 * (size_t)f + 1], a.nb_phase, a.mean);
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        const int q0 = c_gauss_off[gs], ng = c_gauss_off[gs + 1] - q0;
        for (int k = 0; k < ng; k++) {
            const int q = q0 + k;
            const double w = c_gauss[4 * q];
            const double b[3] = {c_gauss[4 * q + 1], c_gauss[4 * q + 2], c_gauss[4 * q + 3]};
            const double I = a.fint[12 * (size_t)f + k];
            const V3 fv = mul(norm, (2 * I - c0 - c1) * (c0 - c1) * w * area);
#pragma unroll
            for (int m = 0; m < 3; m++) { acc[m][0] += fv.x * b[m]; acc[m][1] += fv.y * b[m]; acc[m][2] += fv.z * b[m]; }
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const unsigned row = a.apos[a.mesh.face_nodes[3 * (size_t)f + m]];
            lo = min(lo, row); hi = max(hi, row + 1u);
            double *dst = a.rows + 3 * (size_t)row;
            atomicAdd(dst, acc[m][0]);
            atomicAdd(dst + 1, acc[m][1]);
            atomicAdd(dst + 2, acc[m][2]);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && hi > lo) { atomicMin(loc + 2, lo); atomicMax(loc + 3, hi); }
    comm_grid_barrier(loc);
    // (2) push the touched range (and clear the local rows), then the range and the flag
    lo = *reinterpret_cast<volatile unsigned *>(loc + 2);
    hi = *reinterpret_cast<volatile unsigned *>(loc + 3);
    if (hi < lo) lo = hi = 0u;
    hi = min(hi, c.cap_rows);
    const size_t off = comm_rows_offset(half, c.rank, c.world, c.cap_rows);
    bool wrote = false;
    for (size_t i = 3 * (size_t)lo + gtid; i < 3 * (size_t)hi; i += gsz) {
        const double v = a.rows[i];
        a.rows[i] = 0.0;
        for (int k = 0; k < c.world; k++) reinterpret_cast<double *>(c.win[k] + off)[i] = v;
        wrote = true;
    }
    if (blockIdx.x == 0 && threadIdx.x < c.world) {
        reinterpret_cast<uint2 *>(c.win[threadIdx.x] + kCommRanges)[half * kCommMaxWorld + c.rank] = make_uint2(lo, hi);
        wrote = true;
    }
    if (wrote) __threadfence_system(); // (a system-scope fence per thread is dear: only the threads that stored to a peer)
    __syncthreads();
    // the block that finishes last stores the flags; nobody waits for it here -- every block goes on to wait for the peers
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(loc + 1, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < c.world)
            st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommFlagForce) + half * kCommMaxWorld + c.rank, e);
    }
    // (3) wait for every rank's rows (local polling), add them in rank order
    __shared__ uint2 s_rng[kCommMaxWorld];
    __shared__ int s_failed;
    if (threadIdx.x == 0) s_failed = 0;
    __syncthreads();
    if (threadIdx.x < c.world) {
        if (!comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommFlagForce) + half * kCommMaxWorld + threadIdx.x, e, reinterpret_cast<int *>(mine + kCommError)))
            s_failed = 1;
        const volatile uint2 *rp = reinterpret_cast<const volatile uint2 *>(mine + kCommRanges) + half * kCommMaxWorld + threadIdx.x;
        s_rng[threadIdx.x] = make_uint2(rp->x, rp->y);
    }
    __syncthreads();
    const unsigned n_active = min(*a.n_active_ptr, c.cap_rows);
    const bool failed = s_failed != 0; // a peer never arrived: the rows read as NaN, not as a partial sum
    for (unsigned i = gtid; i < n_active; i += gsz) {
        double x = 0, y = 0, z = 0;
        bool any = false;
        if (failed) { x = y = z = __longlong_as_double(0x7ff8000000000000ll); any = true; }
        else
        for (int k = 0; k < c.world; k++) {
            if (i < s_rng[k].x || i >= s_rng[k].y) continue;
            const double *src = reinterpret_cast<const double *>(mine + comm_rows_offset(half, k, c.world, c.cap_rows)) + 3 * (size_t)i;
            x += __ldcv(src);
            y += __ldcv(src + 1);
            z += __ldcv(src + 2);
            any = true;
        }
        if (any) {
            const size_t node = a.active_nodes[i];
            a.forces[3 * node] = x;
            a.forces[3 * node + 1] = y;
            a.forces[3 * node + 2] = z;
        }
    }
    // scratch of the other half, for the next evaluation (nobody is looking at it now)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned *o = reinterpret_cast<unsigned *>(mine + kCommLocal) + 4 * (half ^ 1);
        o[0] = 0u; o[1] = 0u; o[2] = 0xffffffffu; o[3] = 0u;
    }
}

// ---- sharded snapshot commit ----------------------------------------------------------------------------------------------
// Every rank needs the whole snapshot (positions, tets, labels, interface faces), but the host copy of it is the same on every
// rank: rank r sends only bytes [b_r, e_r) of the flattened snapshot over PCIe, into the snapshot region of its own window, and
// comm_snapshot_push_kernel copies that slice into the same place of every peer's window over NVLink, then its flag;
// comm_snapshot_gather_kernel waits for the slices of all ranks (local polling), copies the assembled region into the snapshot
// buffers and acknowledges the epoch; nobody pushes epoch e + 1 into a window whose owner has not acknowledged e.
__global__ void __launch_bounds__(256) comm_snapshot_push_kernel(CommView c, size_t begin, size_t end)
{
    unsigned char *mine = c.win[c.rank];
    __shared__ unsigned long long s_e;
    if (threadIdx.x == 0) s_e = *reinterpret_cast<const unsigned long long *>(mine + kCommSnapEpoch) + 1; // advanced by the last block
    __syncthreads();
    const unsigned long long e = s_e;
    // the owners of the windows written here must be done with the previous snapshot
    if (threadIdx.x < c.world && threadIdx.x != c.rank)
        comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommAckSnap) + threadIdx.x, e - 1, reinterpret_cast<int *>(mine + kCommError));
    __syncthreads();
    const uint4 *src = reinterpret_cast<const uint4 *>(mine + c.snap_off);
    const size_t i0 = begin / 16, i1 = (end + 15) / 16;
    for (size_t i = i0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < i1; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int k = 0; k < c.world; k++)
            if (k != c.rank) reinterpret_cast<uint4 *>(c.win[k] + c.snap_off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned *>(mine + kCommSnapDone), 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) {
            *reinterpret_cast<unsigned *>(mine + kCommSnapDone) = 0u;
            *reinterpret_cast<unsigned long long *>(mine + kCommSnapEpoch) = e;
        }
        __threadfence_system();
        if (threadIdx.x < c.world) st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommFlagSnap) + c.rank, e);
    }
}

struct SnapshotParts {
    unsigned char *dst[6]; // positions, tet nodes, tet labels, face nodes, face tets, face flags
    size_t off[6], len[6]; // place in the snapshot region (256-byte aligned), bytes
};

// waits for the slices of all ranks (local polling), copies the assembled region into the snapshot buffers, acknowledges.
// (A kernel, not copy-engine copies: those would queue behind this wait and, with several contexts on one device, in front of
// a peer's upload.)
__global__ void __launch_bounds__(256) comm_snapshot_gather_kernel(CommView c, SnapshotParts p)
{
    unsigned char *mine = c.win[c.rank];
    // the push kernel of this commit has run (stream order): the epoch is the current one
    const unsigned long long e = *reinterpret_cast<const unsigned long long *>(mine + kCommSnapEpoch);
    if (threadIdx.x < c.world) comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommFlagSnap) + threadIdx.x, e, reinterpret_cast<int *>(mine + kCommError));
    __syncthreads();
    const unsigned char *region = mine + c.snap_off;
    for (int k = 0; k < 6; k++) {
        const size_t n16 = p.len[k] / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(region + p.off[k]);
        uint4 *dst = reinterpret_cast<uint4 *>(p.dst[k]);
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldcv(src + i);
        for (size_t i = n16 * 16 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < p.len[k]; i += (size_t)gridDim.x * blockDim.x) p.dst[k][i] = region[p.off[k] + i];
    }
    __threadfence();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned *>(mine + kCommSnapDone), 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) { // every block has read the region: the peers may overwrite it with the next snapshot
        if (threadIdx.x == 0) *reinterpret_cast<unsigned *>(mine + kCommSnapDone) = 0u;
        __threadfence_system();
        if (threadIdx.x < c.world) st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommAckSnap) + c.rank, e);
    }
}

} // namespace dsc
