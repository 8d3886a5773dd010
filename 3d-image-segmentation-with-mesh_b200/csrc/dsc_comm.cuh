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
//   force_scatter_rows_kernel    force_atomic_kernel scattering into compact local rows (one per node with interface faces).
//   comm_force_push_kernel       copies the local rows into this rank's slot of every window and clears them; the last block
//                                to finish stores the flags.
//   comm_force_reduce_kernel     waits for all flags (local polling), then every thread adds one row of all slots in rank
//                                order and stores it into the full force array.
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

struct CommView {
    unsigned char *win[kCommMaxWorld]; // window of every rank as mapped here; win[rank] is this rank's own
    int rank, world;
    unsigned cap_rows;
};

DSC_HD size_t comm_window_bytes(int world, unsigned cap_rows) { return kCommRows + 2 * (size_t)world * 3 * cap_rows * sizeof(double); }
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
                                                                                  double s_sq, double s_vol, double *sums, double *mean)
{
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
    if (FIXED) {
        const long long *partials = static_cast<const long long *>(partials_v);
        long long r = 0;
        for (int i = lane; i < n_rows; i += 32) r += partials[(size_t)i * 3 * kMaxPhase + col];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
        bits = (unsigned long long)r;
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
    if (threadIdx.x < c.world) {
        st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommFlagSums) + half * kCommMaxWorld + c.rank, e);
        comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommFlagSums) + half * kCommMaxWorld + threadIdx.x, e, reinterpret_cast<int *>(mine + kCommError));
    }
    __syncthreads();
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

struct ForceRowsArgs {
    VolumeView vol;
    MeshView mesh;
    int nb_phase;
    const double *mean;
    const unsigned *apos;       // node -> row (exclusive scan of "has interface faces")
    double *rows;               // local rows, 3 per active node, zero on entry
    unsigned begin, end;        // faces of this rank
};

// force_atomic_kernel with compact destination rows
template <typename VoxT> __global__ void __launch_bounds__(128) force_scatter_rows_kernel(const ForceRowsArgs a)
{
    const unsigned f = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.end) return;
    const VoxT *__restrict__ vol = static_cast<const VoxT *>(a.vol.data);
    FaceGeom g;
    unsigned n[3];
    if (!face_setup(a.mesh, f, a.nb_phase, a.mean, g, n)) return;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int q = c_gauss_off[g.gs]; q < c_gauss_off[g.gs + 1]; q++) {
        const double w = c_gauss[4 * q];
        const double b[3] = {c_gauss[4 * q + 1], c_gauss[4 * q + 2], c_gauss[4 * q + 3]};
        const V3 p = add(add(mul(g.p[0], b[0]), mul(g.p[1], b[1])), mul(g.p[2], b[2]));
        const double I = fetch_trilinear(vol, a.vol.X, a.vol.Y, a.vol.Z, p);
        const V3 fv = mul(g.norm, (2 * I - g.c0 - g.c1) * (g.c0 - g.c1) * w * g.area);
#pragma unroll
        for (int k = 0; k < 3; k++) { acc[k][0] += fv.x * b[k]; acc[k][1] += fv.y * b[k]; acc[k][2] += fv.z * b[k]; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double *dst = a.rows + 3 * (size_t)a.apos[n[k]];
        atomicAdd(dst, acc[k][0]);
        atomicAdd(dst + 1, acc[k][1]);
        atomicAdd(dst + 2, acc[k][2]);
    }
}

// local rows -> this rank's slot in every window (and cleared for the next evaluation); the last block stores the flags
__global__ void __launch_bounds__(256) comm_force_push_kernel(CommView c, double *__restrict__ rows, const unsigned *__restrict__ n_active_ptr)
{
    unsigned char *mine = c.win[c.rank];
    const unsigned long long e = *reinterpret_cast<const unsigned long long *>(mine + kCommEpoch);
    const int half = (int)(e & 1);
    const size_t n = 3 * (size_t)min(*n_active_ptr, c.cap_rows);
    const size_t off = comm_rows_offset(half, c.rank, c.world, c.cap_rows);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = rows[i];
        rows[i] = 0.0;
        for (int k = 0; k < c.world; k++) reinterpret_cast<double *>(c.win[k] + off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned *>(mine + kCommDone), 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *reinterpret_cast<unsigned *>(mine + kCommDone) = 0u;
        __threadfence_system();
        if (threadIdx.x < c.world) st_release_sys(reinterpret_cast<unsigned long long *>(c.win[threadIdx.x] + kCommFlagForce) + half * kCommMaxWorld + c.rank, e);
    }
}

// slots of all ranks added in rank order into the full force array (which is zero everywhere else)
__global__ void __launch_bounds__(256) comm_force_reduce_kernel(CommView c, const unsigned *__restrict__ active_nodes, const unsigned *__restrict__ n_active_ptr,
                                                                double *__restrict__ forces)
{
    unsigned char *mine = c.win[c.rank];
    const unsigned long long e = *reinterpret_cast<const unsigned long long *>(mine + kCommEpoch);
    const int half = (int)(e & 1);
    if (threadIdx.x < c.world) comm_wait(reinterpret_cast<const unsigned long long *>(mine + kCommFlagForce) + half * kCommMaxWorld + threadIdx.x, e,
                                         reinterpret_cast<int *>(mine + kCommError));
    __syncthreads();
    const unsigned n_active = min(*n_active_ptr, c.cap_rows);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_active; i += gridDim.x * blockDim.x) {
        double x = 0, y = 0, z = 0;
        for (int k = 0; k < c.world; k++) {
            const double *src = reinterpret_cast<const double *>(mine + comm_rows_offset(half, k, c.world, c.cap_rows)) + 3 * (size_t)i;
            x += __ldcv(src);
            y += __ldcv(src + 1);
            z += __ldcv(src + 2);
        }
        const size_t node = active_nodes[i];
        forces[3 * node] = x;
        forces[3 * node + 1] = y;
        forces[3 * node + 2] = z;
    }
}

} // namespace dsc
