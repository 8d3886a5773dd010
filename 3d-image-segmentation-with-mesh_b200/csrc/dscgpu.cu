// dscgpu.cu -- C ABI (include/dscgpu.h) over the sm_100a kernels in dsc_kernels.cuh.
// Host-side responsibilities: context and HBM buffer management, pinned staging of the mesh
// snapshot, the node->face CSR (counting sort), launch configuration, CUDA-event timers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <chrono>
#include <vector>
#include <algorithm>
#include <new>
#include <cuda.h> // CUtensorMap; the encoder is fetched with cudaGetDriverEntryPoint (no link against libcuda)

#include "../../include/dscgpu.h"
#include "dsc_kernels.cuh"
#include "dsc_tet_staged.cuh"
#include "dsc_comm.cuh"
#include "dsc_tables.h"

using namespace dsc;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256; // headroom: meshes change size every iteration
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

constexpr unsigned kMaxPartialRows = 32768;

size_t dtype_size(int dt) { return dt == DSCGPU_U8 ? 1 : dt == DSCGPU_U16 ? 2 : dt == DSCGPU_F32 ? 4 : 8; }

} // namespace

struct dscgpu_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_chunk[20];
    int pipeline_chunks = 4; // tet-array upload chunks overlapped with the tet kernel in dscgpu_image_force_eval (0 = off)
    double pipeline_ratio = 0.85; // size of a chunk relative to the one before it (bit 2 of pipeline_flags; DSCGPU_PIPELINE_RATIO)
    int pipeline_flags = 7;  // bit 2: shrinking chunks; bit 0: face arrays right behind tet chunk 0 and the node -> face table built while chunk 1 is in flight; bit 1: exact list worked off per chunk (DSCGPU_PIPELINE_FLAGS)
    int dims[3] = {0, 0, 0};
    int dtype = DSCGPU_F32;
    void *d_vol = nullptr;
    double *d_sum_table = nullptr, *d_sq_table = nullptr; // z prefix sums of I (image3d::_sum_table) and of I^2 (energy read-out)
    bool has_sum_table = false, has_sq_table = false;
    double *d_tet_bary = nullptr;
    float4 *d_tet_bary_f = nullptr;
    float4 *d_tet_lat = nullptr;
    int force_thread_gather = 0; // 1: thread-per-node gather (first implementation) instead of warp-per-node
    int hot_want_sq = 0; // dscgpu_image_force_eval / dscgpu_tet_phase_sums_dev skip the second moment unless asked
    int tet_variant = 131072 | 128 | 16384; // see dscgpu_set_tet_variant; default = the staged kernel, w2 (clamp-free full batches on the sums-only path) for volumes it cannot describe
    // staged kernel (variant bit 131072, dsc_tet_staged.cuh): tensor-map family in device memory, tile counter, fixed-point scales
    TmapBlob *d_stg_maps = nullptr;
    unsigned *d_stg_counter = nullptr; // [0] tile counter, [1] length of the exact list, [2] its worked-off part, then (at byte 16) the row of 3 * kMaxPhase fixed-point sums
    DevBuf stg_exact;           // the exact list
    TetArgs stg_args;           // arguments of the staged launches of the running pass (for tet_exact_list_kernel at its end)
    int stg_list_kernel = -1;   // which instantiation works the list off: (want_sq << 1 | per_tet), -1: nothing on the list that a launch has not been issued for
    bool stg_pass_open = false; // staged launches of this pass have been issued and its finalize has not
    bool stg_counters_dirty = true; // tile counter / list length may be non-zero (first use, or a pass that did not reach its finalize)
    int stg_state = 0;          // 0 not tried, 1 ready, -1 unavailable
    int stg_slot_bytes = 24 * 1024, stg_slots = 8;
    double stg_scale[3] = {1.0, 1.0, 1.0};
    bool tet_rows_fixed = false; // the partial rows of the running tet pass are fixed-point integers (staged kernel)
    float vol_max = 1.0f;        // largest voxel of a float volume (volume_flags_kernel)
    bool vol_flags_valid = false, vol_simple = false; // the same property from a scan of the volume (volume_flags_kernel)
    unsigned *d_vol_flags = nullptr;
    // peer-memory exchange (dsc_comm.cuh): this rank's window, the peers' windows as mapped here
    unsigned char *comm_win[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool comm_ipc[8] = {false, false, false, false, false, false, false, false}; // mapped with cudaIpcOpenMemHandle (to be closed)
    int comm_rank = 0, comm_world = 0;
    unsigned comm_cap_rows = 0;
    size_t comm_snap_bytes = 0; // snapshot region of the window (sharded commit), 0: none
    bool comm_connected = false;
    uint32_t upd_n_pos = 0, upd_n_tet = 0, upd_n_faces = 0; // the staged update (dscgpu_update_staging_get)
    bool upd_keys = false, upd_staged = false;
    bool forces_resident = false, iforces_resident = false; // ctx->forces / ctx->iforce (and nflags, vbound) hold the results of the last force calls on this snapshot
    DevBuf ndisp, ndest, crows;
    DevBuf ord_off, ord_idx;     // dscgpu_set_node_face_order: get_faces(node) order of the interface faces, CSR by node key
    bool has_face_order = false;
    int force_node_form = 3;     // 3: force_node3_kernel (lane per component), 2: force_node2_kernel (8 lanes per node through shared memory); DSCGPU_FORCE_NODE_FORM
    int overlap_geom = 0;        // tet passes of the per-iteration entry points start the mean-independent half of the force pass next to their exact-list kernel
    int compact_forces = 0;      // dscgpu_image_force_eval(forces == NULL) reads back the rows of the nodes with interface faces only
    PinBuf h_nact;               // their number, copied to the host when the node -> face table is built
    cudaEvent_t ev_nact = nullptr;
    uint32_t compact_n = 0xffffffffu; // rows of the last compact read-back (0xffffffff: the last read-back was the full array)
    const uint32_t *compact_keys = nullptr; // their node keys (pinned)
    bool init_means_valid = false; // o_mean holds the per-tet means of the current snapshot (dscgpu_init_histogram)
    double *comm_rows = nullptr; // local scatter target of the force exchange: 3 doubles per active node, zero between evaluations

    bool has_snapshot = false;
    uint32_t n_nodes = 0, n_tets = 0, n_faces = 0;
    DevBuf pos4, pos4f, tet_nodes, tet_label, face_nodes, face_tets, face_isb, node_off, node_inc, csr_count, csr_cursor, csr_flag, csr_apos, node_active, fbuf, fset, tet_face_nodes, tet_face_tets;
    bool has_tet_faces = false;
    // pinned staging (host side of the snapshot)
    PinBuf h_pos4, h_tet_nodes, h_tet_label, h_face_nodes, h_face_tets, h_face_isb, h_node_off, h_node_inc, h_node_active, h_out, h_flag;
    bool validation_pending = false;
    uint32_t st_nodes = 0, st_tets = 0, st_faces = 0;
    bool staging_valid = false;

    uint32_t tet_b = 0, tet_e = 0, face_b = 0, face_e = 0, node_b = 0, node_e = 0;
    bool partition_stale = false; // the face list changed size under a partition: the ranges must be set again before the next evaluation

    DevBuf partials, sums, mean, forces, counters;
    DevBuf o_total, o_vol, o_mean, o_energy, o_variation, o_hash, o_nsamp;
    DevBuf fgeo, flab, fint; // mean-independent part of the external force (force_geom_kernel)
    bool geom_async = false;  // force_geom_kernel of the current snapshot is running / has run on copy_stream (join on ev_chunk[19])
    DevBuf ray_count, ray_offset, ray_cursor, hits, block_sums, ray_psum, ray_pcnt, misc, vbound, pit, nflags, iterm, inorm, iskip, iforce, upd;
    PinBuf h_upd;

    cudaEvent_t ev[DSCGPU_T_COUNT][2];
    bool ev_used[DSCGPU_T_COUNT];
    bool timers_on = false;
    // DSCGPU_E2E_TRACE=1: time line of dscgpu_image_force_eval on stderr (events on both streams + when the host issued them)
    bool trace_on = false;
    struct TraceMark { const char *name; cudaEvent_t ev; double host_us; };
    std::vector<TraceMark> trace;
    std::vector<cudaEvent_t> trace_pool;
    std::chrono::steady_clock::time_point trace_t0;
    uint64_t launches = 0;
    int lanes_per_tet = 0;
    int auto_lanes = 8;
    int tet_grid = 0;
    std::string err;
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            char b__[512];                                                                            \
            snprintf(b__, sizeof b__, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                           \
            return DSCGPU_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)

#define ARG(cond, msg)                  \
    do {                                \
        if (!(cond)) {                  \
            ctx->err = msg;             \
            return DSCGPU_ERR_ARG;      \
        }                               \
    } while (0)

#define NEED_SNAPSHOT()                                              \
    do {                                                             \
        if (!ctx) return DSCGPU_ERR_ARG;                             \
        if (!ctx->has_snapshot) {                                    \
            ctx->err = "dscgpu_set_snapshot has not been called";    \
            return DSCGPU_ERR_NO_SNAPSHOT;                           \
        }                                                            \
    } while (0)

#define NEED_PARTITION()                                                                                                              \
    do {                                                                                                                              \
        if (ctx->partition_stale) {                                                                                                   \
            ctx->err = "the interface-face list changed size under a partition: call dscgpu_set_partition before the next evaluation"; \
            return DSCGPU_ERR_ARG;                                                                                                    \
        }                                                                                                                             \
    } while (0)

namespace {

// every synchronising entry point ends here: waits for the stream and raises a deferred snapshot validation error
int sync_and_check(dscgpu_ctx *ctx)
{
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaStreamSynchronize failed: ") + cudaGetErrorString(e);
        return DSCGPU_ERR_CUDA;
    }
    if (ctx->validation_pending) {
        ctx->validation_pending = false;
        if (*ctx->h_flag.as<int>() != 0) {
            ctx->has_snapshot = false;
            ctx->err = "snapshot rejected: tet_nodes/face_nodes hold a node key >= n_nodes_buffer or face_tets a tet index >= n_tets";
            return DSCGPU_ERR_ARG;
        }
    }
    return DSCGPU_OK;
}

void trace_mark(dscgpu_ctx *c, const char *name, cudaStream_t st)
{
    if (!c->trace_on) return;
    if (c->trace.empty()) c->trace_t0 = std::chrono::steady_clock::now();
    cudaEvent_t e;
    if (c->trace.size() < c->trace_pool.size()) e = c->trace_pool[c->trace.size()];
    else { cudaEventCreate(&e); c->trace_pool.push_back(e); }
    cudaEventRecord(e, st);
    c->trace.push_back({name, e, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - c->trace_t0).count()});
}

void trace_dump(dscgpu_ctx *c)
{
    if (!c->trace_on || c->trace.empty()) return;
    const double done = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - c->trace_t0).count();
    fprintf(stderr, "[dscgpu trace] %-22s %10s %10s\n", "mark", "gpu us", "issued us");
    for (const auto &m : c->trace) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->trace[0].ev, m.ev);
        fprintf(stderr, "[dscgpu trace] %-22s %10.1f %10.1f\n", m.name, ms * 1e3, m.host_us);
    }
    fprintf(stderr, "[dscgpu trace] %-22s %10s %10.1f\n", "host: synchronised", "", done);
    c->trace.clear();
}

struct TimerScope {
    dscgpu_ctx *c;
    int which;
    TimerScope(dscgpu_ctx *ctx, int w) : c(ctx), which(w)
    {
        if (c->timers_on) { cudaEventRecord(c->ev[w][0], c->stream); }
    }
    ~TimerScope()
    {
        if (c->timers_on) { cudaEventRecord(c->ev[which][1], c->stream); c->ev_used[which] = true; }
    }
};

VolumeView vol_view(const dscgpu_ctx *c) { return VolumeView{c->d_vol, c->d_sum_table, c->dims[0], c->dims[1], c->dims[2]}; }

MeshView mesh_view(const dscgpu_ctx *c)
{
    MeshView m;
    m.pos4 = c->pos4.as<double4>();
    m.pos4f = c->pos4f.as<float4>();
    m.tet_nodes = c->tet_nodes.as<uint4>();
    m.tet_label = c->tet_label.as<int>();
    m.face_nodes = c->face_nodes.as<unsigned>();
    m.face_tets = c->face_tets.as<unsigned>();
    m.face_is_boundary = c->face_isb.as<unsigned char>();
    m.node_off = c->node_off.as<unsigned>();
    m.node_inc = c->node_inc.as<unsigned>();
    m.n_nodes = c->n_nodes; m.n_tets = c->n_tets; m.n_faces = c->n_faces;
    return m;
}

// counters: [0] samples, [1] gauss points, [2] overflow flag (int), [3] scan total (unsigned)
unsigned long long *ctr(dscgpu_ctx *c, int i) { return c->counters.as<unsigned long long>() + i; }

template <typename VoxT, int G, bool WITH_C> int launch_tet_t(dscgpu_ctx *ctx, TetArgs &a)
{
    auto kern = tet_integrate_kernel<VoxT, G, WITH_C>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) occ = 1;
    unsigned n_work = a.tet_end - a.tet_begin;
    unsigned groups_per_block = 256 / G;
    unsigned need = (n_work + groups_per_block - 1) / groups_per_block;
    unsigned grid = (unsigned)(ctx->num_sms * occ);
    if (need < grid) grid = need ? need : 1;
    if (ctx->tet_grid + grid > kMaxPartialRows) { ctx->err = "too many partial rows"; return DSCGPU_ERR_ARG; }
    a.partials = ctx->partials.as<double>() + (size_t)ctx->tet_grid * 3 * kMaxPhase;
    ctx->tet_grid += (int)grid;
    kern<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

// float volumes: one pass that tells whether every voxel is a non-negative finite number (scaled widening, dsc_kernels.cuh)
int ensure_volume_flags(dscgpu_ctx *ctx)
{
    if (ctx->vol_flags_valid) return DSCGPU_OK;
    ctx->vol_simple = false;
    if (ctx->dtype == DSCGPU_F32) {
        if (!ctx->d_vol_flags) { CK(cudaMalloc(&ctx->d_vol_flags, 4 * sizeof(unsigned))); CK(cudaMemset(ctx->d_vol_flags, 0, 4 * sizeof(unsigned))); }
        CK(cudaMemsetAsync(ctx->d_vol_flags, 0, 2 * sizeof(unsigned), ctx->stream));
        const size_t n = (size_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
        volume_flags_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(static_cast<const float *>(ctx->d_vol), n, ctx->d_vol_flags);
        ctx->launches++;
        CK(cudaGetLastError());
        unsigned flags[2] = {1, 0};
        CK(cudaMemcpyAsync(flags, ctx->d_vol_flags, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->vol_simple = (flags[0] & 1u) == 0;
        memcpy(&ctx->vol_max, &flags[1], sizeof(float)); // bit pattern order = value order for non-negative floats
    }
    ctx->vol_flags_valid = true;
    return DSCGPU_OK;
}

// Staged kernel: one tensor map per (row bytes, rows) pair of a z slice of a voxel box -- 16 x 16 maps of 128 bytes in device
// memory, picked per tile by the kernel -- plus the tile counter and the fixed-point scales of the per-label sums.
// Unavailable (the w2 kernel runs instead): row pitch of the volume not a multiple of 16 bytes, FP64 volumes, float volumes with
// negative / non-finite voxels, no driver entry point.
int ensure_staged(dscgpu_ctx *ctx)
{
    if (ctx->stg_state != 0) return DSCGPU_OK;
    ctx->stg_state = -1;
    const size_t es = dtype_size(ctx->dtype);
    const cuuint64_t X = (cuuint64_t)ctx->dims[0], Y = (cuuint64_t)ctx->dims[1], Z = (cuuint64_t)ctx->dims[2];
    if (ctx->dtype == DSCGPU_F64 || (X * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(ctx->d_vol) & 15) != 0) return DSCGPU_OK;
    double vmax = 1.0;
    if (ctx->dtype == DSCGPU_F32) {
        int rf = ensure_volume_flags(ctx);
        if (rf) return rf;
        if (!ctx->vol_simple) return DSCGPU_OK;
        vmax = (double)ctx->vol_max;
    }
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return DSCGPU_OK;
    }
    const CUtensorMapDataType dt = ctx->dtype == DSCGPU_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : ctx->dtype == DSCGPU_U16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    std::vector<TmapBlob> maps((size_t)kStgMaxUnits * kStgMaxRows * kStgMaxSlices);
    const cuuint64_t gdim[3] = {X, Y, Z}, gstr[2] = {X * es, X * Y * es};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int u = 0; u < kStgMaxUnits; u++)
        for (int r = 0; r < kStgMaxRows; r++)
            for (int sl = 0; sl < kStgMaxSlices; sl++) {
                alignas(64) CUtensorMap tm;
                const cuuint32_t box[3] = {(cuuint32_t)((u + 1) * 16 / es), (cuuint32_t)(r + 1), (cuuint32_t)(sl + 1)};
                if (reinterpret_cast<encode_fn>(fn)(&tm, dt, 3, ctx->d_vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return DSCGPU_OK;
                memcpy(&maps[((size_t)u * kStgMaxRows + r) * kStgMaxSlices + sl], &tm, sizeof(CUtensorMap));
            }
    if (!ctx->d_stg_maps) CK(cudaMalloc(&ctx->d_stg_maps, maps.size() * sizeof(TmapBlob)));
    if (!ctx->d_stg_counter) CK(cudaMalloc(&ctx->d_stg_counter, 16 + 3 * kMaxPhase * sizeof(unsigned long long)));
    CK(cudaMemcpy(ctx->d_stg_maps, maps.data(), maps.size() * sizeof(TmapBlob), cudaMemcpyHostToDevice));
    // fixed point: per-label sums stay below 64 x (voxels of the volume) x (largest intensity) -- tets cover the volume once, a
    // DSC mesh pads it by one edge -- and must stay below 2^62; tet_finalize_fixed_kernel reports a sum that did not
    const double nvox = (double)X * (double)Y * (double)Z;
    auto scale_for = [](double bound) { int e = 0; frexp(bound, &e); int k = 62 - e; if (k > 60) k = 60; if (k < -60) k = -60; return ldexp(1.0, k); };
    const double vm = vmax > 1.0e-30 ? vmax : 1.0e-30;
    ctx->stg_scale[0] = scale_for(64.0 * nvox * vm);
    ctx->stg_scale[1] = scale_for(64.0 * nvox * vm * vm);
    ctx->stg_scale[2] = scale_for(64.0 * nvox);
    ctx->stg_state = 1;
    return DSCGPU_OK;
}

template <typename VoxT, bool WANT_SQ, bool PER_TET, bool F32S, int PMAX> int launch_staged_t(dscgpu_ctx *ctx, TetArgs &a)
{
    auto kern = tet_staged_kernel<VoxT, WANT_SQ, PER_TET, F32S, PMAX>;
    const int smem = ctx->stg_slots * (ctx->stg_slot_bytes + kStgRecBytes);
    static bool attr_set = false; // per instantiation
    static int attr_dev = -1;
    if (!attr_set || attr_dev != ctx->device) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        attr_set = true; attr_dev = ctx->device;
    }
    const unsigned n_work = a.tet_end - a.tet_begin;
    const unsigned n_tiles = (n_work + 31) / 32;
    unsigned grid = (unsigned)ctx->num_sms;
    const unsigned need = (n_tiles + kStgProd - 1) / kStgProd;
    if (need < grid) grid = need ? need : 1;
    // (the list buffer is sized with the snapshot: it cannot grow while the stream is being captured)
    if (ctx->stg_exact.cap < (size_t)ctx->n_tets * sizeof(unsigned)) { ctx->err = "exact-list buffer smaller than the snapshot"; return DSCGPU_ERR_ARG; }
    a.stg_acc = reinterpret_cast<unsigned long long *>(ctx->d_stg_counter + 4);
    ctx->tet_rows_fixed = true;
    a.stg_maps = ctx->d_stg_maps;
    a.stg_slot_bytes = ctx->stg_slot_bytes;
    a.stg_slots = ctx->stg_slots;
    a.stg_counter = ctx->d_stg_counter;
    a.stg_exact_count = ctx->d_stg_counter + 1;
    a.stg_exact_done = ctx->d_stg_counter + 2;
    a.stg_exact_list = ctx->stg_exact.as<unsigned>();
    // tiles per claim: a chunk is also the grain of the tail, so launches with few tiles per producer warp claim them one at a time
    const unsigned per_warp = n_tiles / (grid * kStgProd);
    a.stg_chunk = per_warp >= 64 ? kStgMaxChunk : (per_warp >= 24 ? 2 : 1);
    for (int k = 0; k < 3; k++) a.stg_scale[k] = ctx->stg_scale[k];
    // the counters are zero after the finalize of the previous pass; a second range launch of one pass restarts the tile counter only
    if (ctx->stg_counters_dirty) CK(cudaMemsetAsync(ctx->d_stg_counter, 0, 16 + 3 * kMaxPhase * sizeof(unsigned long long), ctx->stream));
    else if (ctx->stg_pass_open) CK(cudaMemsetAsync(ctx->d_stg_counter, 0, sizeof(unsigned), ctx->stream));
    ctx->stg_counters_dirty = false;
    ctx->stg_pass_open = true;
    kern<<<grid, (kStgProd + kStgCons) * 32, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->stg_args = a;
    ctx->stg_list_kernel = (WANT_SQ ? 2 : 0) | (PER_TET ? 1 : 0);
    return DSCGPU_OK;
}

// end of a pass with staged launches: the tets they left to the exact path, one launch for the whole pass
template <typename VoxT, bool F32S> int launch_exact_list_v(dscgpu_ctx *ctx)
{
    TetArgs &a = ctx->stg_args;
    // worked off by warps, one tet each: enough of them to hide the dependent loads, no more than the list can be expected to feed
    const unsigned n_work = ctx->tet_e - ctx->tet_b;
    unsigned grid2 = (unsigned)ctx->num_sms * 4;
    const unsigned need2 = (n_work / 64 + 7) / 8;
    if (need2 < grid2) grid2 = need2 ? need2 : 1;
    switch (ctx->stg_list_kernel) {
    case 0: tet_exact_list_kernel<VoxT, false, false, F32S><<<grid2, 256, 0, ctx->stream>>>(a); break;
    case 1: tet_exact_list_kernel<VoxT, false, true, F32S><<<grid2, 256, 0, ctx->stream>>>(a); break;
    case 2: tet_exact_list_kernel<VoxT, true, false, F32S><<<grid2, 256, 0, ctx->stream>>>(a); break;
    default: tet_exact_list_kernel<VoxT, true, true, F32S><<<grid2, 256, 0, ctx->stream>>>(a); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->stg_list_kernel = -1;
    return DSCGPU_OK;
}

// more_follow: further staged launches of the same pass come after this one (the chunks of the pipelined upload): the part of the list
// worked off so far is remembered on the device, so that every entry is evaluated exactly once
int launch_exact_list(dscgpu_ctx *ctx, bool more_follow = false)
{
    if (ctx->stg_list_kernel < 0) return DSCGPU_OK;
    int r;
    switch (ctx->dtype) {
    case DSCGPU_U8: r = launch_exact_list_v<unsigned char, false>(ctx); break;
    case DSCGPU_U16: r = launch_exact_list_v<unsigned short, false>(ctx); break;
    default: r = launch_exact_list_v<float, true>(ctx); break;
    }
    if (r == DSCGPU_OK && more_follow) {
        copy_u32_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_stg_counter + 2, ctx->d_stg_counter + 1);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return r;
}

template <typename VoxT, bool F32S, int PMAX> int launch_staged_p(dscgpu_ctx *ctx, TetArgs &a)
{
    const bool per_tet = a.tet_total || a.tet_vol || a.tet_mean;
    if (a.want_sq) return per_tet ? launch_staged_t<VoxT, true, true, F32S, PMAX>(ctx, a) : launch_staged_t<VoxT, true, false, F32S, PMAX>(ctx, a);
    return per_tet ? launch_staged_t<VoxT, false, true, F32S, PMAX>(ctx, a) : launch_staged_t<VoxT, false, false, F32S, PMAX>(ctx, a);
}

// per-label accumulators live in registers of the consumer lanes: four labels cover the usual cases with half the registers of eight
template <typename VoxT, bool F32S> int launch_staged_v(dscgpu_ctx *ctx, TetArgs &a)
{
    return a.nb_phase <= 4 ? launch_staged_p<VoxT, F32S, 4>(ctx, a) : launch_staged_p<VoxT, F32S, kMaxPhase>(ctx, a);
}

int launch_staged(dscgpu_ctx *ctx, TetArgs &a)
{
    switch (ctx->dtype) {
    case DSCGPU_U8: return launch_staged_v<unsigned char, false>(ctx, a);
    case DSCGPU_U16: return launch_staged_v<unsigned short, false>(ctx, a);
    default: return launch_staged_v<float, true>(ctx, a);
    }
}

template <typename VoxT, bool WANT_SQ, bool PER_TET, bool BRICK = false, bool SIMPLE = false, bool PACK = false, int LEAN = 0, bool PF = false> int launch_w2_t(dscgpu_ctx *ctx, TetArgs &a)
{
    auto kern = tet_integrate_w2_kernel<VoxT, WANT_SQ, PER_TET, BRICK, SIMPLE, PACK, LEAN, PF>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) occ = 1;
    unsigned n_work = a.tet_end - a.tet_begin;
    unsigned need = (n_work + 255) / 256;
    unsigned grid = (unsigned)(ctx->num_sms * occ);
    if (need < grid) grid = need ? need : 1;
    if (ctx->tet_grid + grid > kMaxPartialRows) { ctx->err = "too many partial rows"; return DSCGPU_ERR_ARG; }
    a.partials = ctx->partials.as<double>() + (size_t)ctx->tet_grid * 3 * kMaxPhase;
    ctx->tet_grid += (int)grid;
    kern<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

template <typename VoxT> int launch_w2(dscgpu_ctx *ctx, TetArgs &a)
{
    const bool per_tet = a.tet_total || a.tet_vol || a.tet_mean;
    if (a.lean && !a.want_sq && !per_tet) { // sums only: full batches without row clamps and masks
        if constexpr (sizeof(VoxT) == 4 && !VoxTraits<VoxT>::integral) {
            if (a.vol_simple) return launch_w2_t<VoxT, false, false, false, true, false, 1>(ctx, a);
        }
        return launch_w2_t<VoxT, false, false, false, false, false, 1>(ctx, a);
    }
    if constexpr (sizeof(VoxT) == 4 && !VoxTraits<VoxT>::integral) {
        if (a.vol_simple) { // non-negative finite floats: scaled widening
            if (a.want_sq) return per_tet ? launch_w2_t<VoxT, true, true, false, true>(ctx, a) : launch_w2_t<VoxT, true, false, false, true>(ctx, a);
            return per_tet ? launch_w2_t<VoxT, false, true, false, true>(ctx, a) : launch_w2_t<VoxT, false, false, false, true>(ctx, a);
        }
    }
    if (a.want_sq) return per_tet ? launch_w2_t<VoxT, true, true>(ctx, a) : launch_w2_t<VoxT, true, false>(ctx, a);
    return per_tet ? launch_w2_t<VoxT, false, true>(ctx, a) : launch_w2_t<VoxT, false, false>(ctx, a);
}

template <typename VoxT> int launch_tet_v(dscgpu_ctx *ctx, TetArgs &a, int G, bool with_c)
{
    if (G == 32) return with_c ? launch_tet_t<VoxT, 32, true>(ctx, a) : launch_tet_t<VoxT, 32, false>(ctx, a);
    return with_c ? launch_tet_t<VoxT, 8, true>(ctx, a) : launch_tet_t<VoxT, 8, false>(ctx, a);
}


// Which kernel runs a tet pass (dscgpu_set_tet_variant): bit 131072 the staged kernel (dsc_tet_staged.cuh), bit 128 the w2 kernel
// (bit 16384: clamp-free full batches on the sums-only path; bit 4096: general float widening), otherwise -- and whenever candidate
// values c are given -- the plain G-lanes-per-tet kernel, which is the reference's sequence verbatim.
int launch_tet(dscgpu_ctx *ctx, TetArgs &a, bool with_c)
{
    int G = ctx->lanes_per_tet ? ctx->lanes_per_tet : ctx->auto_lanes;
    const bool fits32 = (double)ctx->dims[0] * ctx->dims[1] * ctx->dims[2] < 2147483648.0;
    if ((ctx->tet_variant & 131072) && !with_c && fits32) { // staged kernel: lane per tet, voxel boxes in shared memory by TMA
        int rs = ensure_staged(ctx);
        if (rs) return rs;
        if (ctx->stg_state == 1 && (ctx->tet_grid == 0 || ctx->tet_rows_fixed)) return launch_staged(ctx, a); // (a pass is all double rows or all fixed point)
    }
    if ((ctx->tet_variant & 128) && !with_c && fits32) { // warp-level two-phase kernel with lattice coordinates
        a.lean = (ctx->tet_variant & 16384) ? 1 : 0;
        if (!(ctx->tet_variant & 4096)) { // bit 12: keep the general float widening (A/B switch)
            int rf = ensure_volume_flags(ctx);
            if (rf) return rf;
            a.vol_simple = ctx->vol_simple ? 1 : 0;
        }
        switch (ctx->dtype) {
        case DSCGPU_U8: return launch_w2<unsigned char>(ctx, a);
        case DSCGPU_U16: return launch_w2<unsigned short>(ctx, a);
        case DSCGPU_F32: return launch_w2<float>(ctx, a);
        default: return launch_w2<double>(ctx, a);
        }
    }
    if (G != 32) G = 8;
    switch (ctx->dtype) {
    case DSCGPU_U8: return launch_tet_v<unsigned char>(ctx, a, G, with_c);
    case DSCGPU_U16: return launch_tet_v<unsigned short>(ctx, a, G, with_c);
    case DSCGPU_F32: return launch_tet_v<float>(ctx, a, G, with_c);
    default: return launch_tet_v<double>(ctx, a, G, with_c);
    }
}

struct TetPassOut {
    double *d_total = nullptr, *d_vol = nullptr, *d_mean = nullptr, *d_energy = nullptr, *d_variation = nullptr;
};

// launches the tet kernel on [begin, end); partial rows accumulate in ctx->partials until tet_pass_finish
int tet_pass_range(dscgpu_ctx *ctx, int nb_phase, int include_bound, int nc, const double *c, const TetPassOut &o, unsigned begin, unsigned end, int want_sq)
{
    if (end <= begin) return DSCGPU_OK;
    TetArgs a;
    a.vol = vol_view(ctx);
    a.mesh = mesh_view(ctx);
    a.tet_bary = ctx->d_tet_bary;
    a.tet_bary_f = ctx->d_tet_bary_f;
    a.tet_lat = ctx->d_tet_lat;
    a.want_sq = want_sq;
    a.tet_begin = begin; a.tet_end = end;
    a.include_bound = include_bound;
    a.nb_phase = nb_phase;
    a.nc = nc;
    for (int k = 0; k < kMaxPhase; k++) a.c[k] = (c && k < nc) ? c[k] : 0.0;
    a.tet_total = o.d_total; a.tet_vol = o.d_vol; a.tet_mean = o.d_mean; a.tet_energy = o.d_energy; a.tet_variation = o.d_variation;
    a.sample_count = ctr(ctx, 0);
    a.vol_simple = 0;
    a.lean = 0;
    return launch_tet(ctx, a, nc > 0 && (o.d_energy || o.d_variation));
}

int tet_pass_begin(dscgpu_ctx *ctx)
{
    ctx->tet_grid = 0;
    ctx->tet_rows_fixed = false;
    if (ctx->stg_pass_open) { ctx->stg_list_kernel = -1; ctx->stg_pass_open = false; ctx->stg_counters_dirty = true; } // a pass that never reached its finalize
    CK(cudaMemsetAsync(ctr(ctx, 0), 0, sizeof(unsigned long long), ctx->stream));
    return DSCGPU_OK;
}

int tet_pass_finish(dscgpu_ctx *ctx, int nb_phase, double *d_sums_out, double *d_mean_out)
{
    int rl = launch_exact_list(ctx); // (no-op when the caller has launched it inside its timer scope)
    if (rl) return rl;
    TimerScope ts(ctx, DSCGPU_T_PHASE_FINAL);
    if (ctx->tet_rows_fixed)
        tet_finalize_fixed_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<unsigned long long *>(ctx->d_stg_counter + 4), nb_phase, ctx->stg_scale[0], ctx->stg_scale[1],
                                                             ctx->stg_scale[2], d_sums_out, d_mean_out, ctx->d_vol_flags ? ctx->d_vol_flags + 2 : nullptr, ctx->d_stg_counter);
    else
        tet_finalize_kernel<<<1, 32 * 3 * kMaxPhase, 0, ctx->stream>>>(ctx->partials.as<double>(), ctx->tet_grid, nb_phase, d_sums_out, d_mean_out);
    ctx->launches++;
    ctx->stg_pass_open = false; // (both finalize kernels leave the staged kernel's counters zero)
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

// tet pass + finalize over this context's partition; leaves sums (3*nb) and means on the device
int run_tet_pass(dscgpu_ctx *ctx, int nb_phase, int include_bound, int nc, const double *c, double *d_total, double *d_vol, double *d_mean,
                 double *d_energy, double *d_variation, double *d_sums_out, double *d_mean_out, int want_sq = 1, bool fork_geom = false)
{
    TetPassOut o;
    o.d_total = d_total; o.d_vol = d_vol; o.d_mean = d_mean; o.d_energy = d_energy; o.d_variation = d_variation;
    int r = tet_pass_begin(ctx);
    if (r) return r;
    {
        TimerScope ts(ctx, DSCGPU_T_TET);
        r = tet_pass_range(ctx, nb_phase, include_bound, nc, c, o, ctx->tet_b, ctx->tet_e, want_sq);
        if (r) return r;
        // The staged kernel owns every SM (one CTA each, the whole register file); what follows it -- the exact list, the finalize -- is
        // small: the geometry half of the force pass (normals, areas, trilinear samples; it does not need the means) runs next to them
        if (fork_geom && ctx->overlap_geom && ctx->n_faces && !ctx->geom_async) {
            r = dscgpu_face_geometry_async(ctx);
            if (r) return r;
        }
        r = launch_exact_list(ctx);
        if (r) return r;
    }
    return tet_pass_finish(ctx, nb_phase, d_sums_out, d_mean_out);
}

template <typename VoxT> int launch_force_t(dscgpu_ctx *ctx, ForceArgs &a, int mode)
{
    unsigned n = a.end - a.begin;
    if (n == 0) return DSCGPU_OK;
    unsigned grid = (n + 127) / 128;
    if (mode == DSCGPU_FORCE_GATHER) force_gather_kernel<VoxT><<<grid, 128, 0, ctx->stream>>>(a);
    else force_atomic_kernel<VoxT><<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

// mean-independent half of the external force (force_geom_kernel) of this context's face range on stream st
int launch_force_geom(dscgpu_ctx *ctx, cudaStream_t st)
{
    const unsigned nf = ctx->n_faces;
    CK(ctx->fgeo.ensure((size_t)nf * 4 * sizeof(double) + 8));
    CK(ctx->flab.ensure((size_t)nf * 2 * sizeof(int) + 8));
    CK(ctx->fint.ensure((size_t)nf * 12 * sizeof(double) + 8));
    CK(ctx->fset.ensure((size_t)nf + 8));
    ForceGeomArgs fa;
    fa.vol = vol_view(ctx); fa.mesh = mesh_view(ctx);
    fa.fgeo = ctx->fgeo.as<double>(); fa.flab = ctx->flab.as<int>(); fa.fint = ctx->fint.as<double>(); fa.fset = ctx->fset.as<signed char>();
    // multi-GPU: this context evaluates the faces of its partition only; faces outside keep set -1 and contribute on another rank
    const bool whole_faces = ctx->face_b == 0 && ctx->face_e == nf;
    fa.begin = ctx->face_b; fa.end = ctx->face_e;
    if (!whole_faces && nf) CK(cudaMemsetAsync(ctx->fset.p, 0xFF, nf, st));
    if (fa.end > fa.begin) {
        const unsigned grid = (unsigned)(((size_t)(fa.end - fa.begin) * kFaceLanes + 127) / 128);
        switch (ctx->dtype) {
        case DSCGPU_U8: force_geom_kernel<unsigned char><<<grid, 128, 0, st>>>(fa); break;
        case DSCGPU_U16: force_geom_kernel<unsigned short><<<grid, 128, 0, st>>>(fa); break;
        case DSCGPU_F32: force_geom_kernel<float><<<grid, 128, 0, st>>>(fa); break;
        default: force_geom_kernel<double><<<grid, 128, 0, st>>>(fa); break;
        }
        ctx->launches++;
    }
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

int run_forces(dscgpu_ctx *ctx, int nb_phase, const double *d_mean, int mode, double *d_forces)
{
    ForceArgs a;
    a.vol = vol_view(ctx);
    a.mesh = mesh_view(ctx);
    a.nb_phase = nb_phase;
    a.mean = d_mean;
    a.forces = d_forces;
    a.gauss_count = ctr(ctx, 1);
    CK(cudaMemsetAsync(ctr(ctx, 1), 0, sizeof(unsigned long long), ctx->stream));
    TimerScope ts(ctx, DSCGPU_T_FORCE);
    const bool whole_nodes = ctx->node_b == 0 && ctx->node_e == ctx->n_nodes;
    if (mode == DSCGPU_FORCE_GATHER && !ctx->force_thread_gather) {
        // two kernels: per-face geometry and samples (mean-independent: it may already be running on the side stream, see
        // dscgpu_face_geometry_async), then per-node ordered sums with the means
        const unsigned nf = ctx->n_faces;
        if (ctx->geom_async) {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[19], 0));
            ctx->geom_async = false;
        } else {
            int rg = launch_force_geom(ctx, ctx->stream);
            if (rg) return rg;
        }
        const bool whole_faces = ctx->face_b == 0 && ctx->face_e == nf;
        ForceNode2Args na;
        na.mesh = a.mesh; na.fgeo = ctx->fgeo.as<double>(); na.flab = ctx->flab.as<int>(); na.fint = ctx->fint.as<double>();
        na.fset = ctx->fset.as<signed char>(); na.nb_phase = nb_phase; na.mean = d_mean; na.forces = d_forces;
        // with a face partition every node may receive a contribution from this rank; with all faces present the node
        // range of the partition is enough (and the result is the reference's, bit for bit)
        if (whole_faces) {
            na.begin = ctx->node_b; na.end = ctx->node_e;
        } else {
            na.begin = 0; na.end = ctx->n_nodes;
        }
        na.active_nodes = ctx->node_active.as<unsigned>();
        na.n_active = reinterpret_cast<const unsigned *>(ctr(ctx, 5));
        // the node pass only writes nodes that have interface faces: everything else must read as zero
        CK(cudaMemsetAsync(d_forces, 0, sizeof(double) * 3 * (size_t)ctx->n_nodes, ctx->stream));
        const unsigned max_active = std::min<unsigned>(ctx->n_nodes, 3u * nf);
        if (na.end > na.begin && max_active) {
            if (ctx->force_node_form == 2) {
                const unsigned grid = std::min<unsigned>((max_active + 15) / 16, (unsigned)ctx->num_sms * 16);
                force_node2_kernel<<<grid, 128, 0, ctx->stream>>>(na);
            } else { // lane per component: every node resident at once
                const unsigned grid = std::min<unsigned>((max_active + 31) / 32, (unsigned)ctx->num_sms * 16);
                force_node3_kernel<<<grid, 128, 0, ctx->stream>>>(na);
            }
            ctx->launches++;
        }
        CK(cudaGetLastError());
        return DSCGPU_OK;
    }
    if (mode == DSCGPU_FORCE_GATHER) {
        a.begin = ctx->node_b; a.end = ctx->node_e;
        // nodes outside this context's range must read as zero for the all-reduce
        if (!whole_nodes) CK(cudaMemsetAsync(d_forces, 0, sizeof(double) * 3 * (size_t)ctx->n_nodes, ctx->stream));
    } else {
        a.begin = ctx->face_b; a.end = ctx->face_e;
        CK(cudaMemsetAsync(d_forces, 0, sizeof(double) * 3 * (size_t)ctx->n_nodes, ctx->stream));
    }
    switch (ctx->dtype) {
    case DSCGPU_U8: return launch_force_t<unsigned char>(ctx, a, mode);
    case DSCGPU_U16: return launch_force_t<unsigned short>(ctx, a, mode);
    case DSCGPU_F32: return launch_force_t<float>(ctx, a, mode);
    default: return launch_force_t<double>(ctx, a, mode);
    }
}

int exclusive_scan(dscgpu_ctx *ctx, const unsigned *in, unsigned *out, size_t n, unsigned *d_total)
{
    unsigned n_blocks = (unsigned)((n + kScanItems - 1) / kScanItems);
    CK(ctx->block_sums.ensure((size_t)n_blocks * sizeof(unsigned)));
    scan_block_kernel<<<n_blocks, 1024, 0, ctx->stream>>>(in, out, ctx->block_sums.as<unsigned>(), n);
    scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->block_sums.as<unsigned>(), n_blocks, d_total);
    scan_add_kernel<<<n_blocks, 1024, 0, ctx->stream>>>(out, ctx->block_sums.as<unsigned>(), n);
    ctx->launches += 3;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

template <typename VoxT> void launch_resolve_t(dscgpu_ctx *ctx, const ResolveArgs &a, dim3 grid)
{
    zray_resolve_kernel<VoxT><<<grid, 128, 0, ctx->stream>>>(a);
}

// z-ray pass; mode 0 leaves sums/means on the device, mode 1 leaves the data term in sums[0..nb)
int run_zray(dscgpu_ctx *ctx, int nb_phase, int mode, const double *c_host, double *d_sums_out, double *d_mean_out)
{
    const size_t XY = (size_t)ctx->dims[0] * ctx->dims[1];
    const size_t n_rays = XY * nb_phase;
    TimerScope ts(ctx, DSCGPU_T_ZRAY);
    CK(ctx->ray_count.ensure(n_rays * sizeof(unsigned)));
    CK(ctx->ray_offset.ensure(n_rays * sizeof(unsigned)));
    CK(ctx->ray_cursor.ensure(n_rays * sizeof(unsigned)));
    CK(cudaMemsetAsync(ctx->ray_count.p, 0, n_rays * sizeof(unsigned), ctx->stream));
    CK(cudaMemsetAsync(ctx->ray_cursor.p, 0, n_rays * sizeof(unsigned), ctx->stream));
    CK(cudaMemsetAsync(ctr(ctx, 2), 0, 2 * sizeof(unsigned long long), ctx->stream));
    RayArgs ra;
    ra.vol = vol_view(ctx);
    ra.mesh = mesh_view(ctx);
    ra.nb_phase = nb_phase;
    ra.face_begin = 0; ra.face_end = ctx->n_faces; // every context rasterises all faces; rays are partitioned
    ra.ray_count = ctx->ray_count.as<unsigned>();
    ra.ray_offset = ctx->ray_offset.as<unsigned>();
    ra.ray_cursor = ctx->ray_cursor.as<unsigned>();
    ra.hits = nullptr; ra.hits_cap = 0; ra.overflow = nullptr;
    unsigned fgrid = (ctx->n_faces + 127) / 128;
    if (fgrid) {
        zray_raster_kernel<false><<<fgrid, 128, 0, ctx->stream>>>(ra);
        ctx->launches++;
    }
    unsigned *d_total = reinterpret_cast<unsigned *>(ctr(ctx, 3));
    int r = exclusive_scan(ctx, ra.ray_count, ra.ray_offset, n_rays, d_total);
    if (r) return r;
    // The hit buffer keeps its size from pass to pass: only a pass that finds it too small (the first one; a mesh whose interface has
    // grown by more than the headroom) reads the total back and repeats -- check_overflow does that after the fact -- so the steady
    // state has no host round trip in the middle of the pass and can be captured in a CUDA graph.
    if (ctx->hits.cap == 0) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
            unsigned total_hits = 0;
            CK(cudaMemcpyAsync(&total_hits, d_total, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(ctx->hits.ensure(((size_t)total_hits + 1) * sizeof(uint2)));
        }
    }
    ra.hits = ctx->hits.as<uint2>();
    ra.hits_cap = (unsigned)std::min<size_t>(ctx->hits.cap / sizeof(uint2), 0xffffffffu);
    ra.overflow = reinterpret_cast<int *>(ctr(ctx, 2));
    if (fgrid) {
        zray_raster_kernel<true><<<fgrid, 128, 0, ctx->stream>>>(ra);
        ctx->launches++;
    }
    ResolveArgs rs;
    rs.vol = vol_view(ctx);
    rs.nb_phase = nb_phase;
    rs.mode = mode;
    for (int k = 0; k < kMaxPhase; k++) rs.c[k] = (c_host && k < nb_phase) ? c_host[k] : 0.0;
    rs.ray_count = ra.ray_count; rs.ray_offset = ra.ray_offset; rs.hits = ra.hits; rs.hits_cap = ra.hits_cap;
    rs.sq_table = (mode == 1 && ctx->has_sum_table && ctx->has_sq_table) ? ctx->d_sq_table : nullptr;
    // ray partition: contiguous column range derived from the tet partition fraction
    unsigned col_b = 0, col_e = (unsigned)XY;
    if (ctx->n_tets && !(ctx->tet_b == 0 && ctx->tet_e == ctx->n_tets)) {
        col_b = (unsigned)((double)ctx->tet_b / ctx->n_tets * XY);
        col_e = ctx->tet_e == ctx->n_tets ? (unsigned)XY : (unsigned)((double)ctx->tet_e / ctx->n_tets * XY);
    }
    rs.col_begin = col_b; rs.col_end = col_e;
    unsigned n_blocks = (col_e - col_b + 127) / 128;
    if (n_blocks == 0) n_blocks = 1;
    CK(ctx->ray_psum.ensure((size_t)n_blocks * nb_phase * sizeof(double)));
    CK(ctx->ray_pcnt.ensure((size_t)n_blocks * nb_phase * sizeof(long long)));
    rs.partial_sum = ctx->ray_psum.as<double>();
    rs.partial_cnt = ctx->ray_pcnt.as<long long>();
    rs.overflow = reinterpret_cast<int *>(ctr(ctx, 2));
    dim3 grid(n_blocks, nb_phase);
    switch (ctx->dtype) {
    case DSCGPU_U8: launch_resolve_t<unsigned char>(ctx, rs, grid); break;
    case DSCGPU_U16: launch_resolve_t<unsigned short>(ctx, rs, grid); break;
    case DSCGPU_F32: launch_resolve_t<float>(ctx, rs, grid); break;
    default: launch_resolve_t<double>(ctx, rs, grid); break;
    }
    zray_finalize_kernel<<<1, 32 * kMaxPhase, 0, ctx->stream>>>(rs.partial_sum, rs.partial_cnt, (int)n_blocks, nb_phase, d_sums_out, d_mean_out);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

// end of a z-ray pass (synchronises): a hit buffer that was too small is grown to the size the pass has counted; the caller repeats
// the pass when *again is set
int check_overflow(dscgpu_ctx *ctx, bool *again = nullptr)
{
    unsigned of[4] = {0, 0, 0, 0}; // counters [2] (flags) and [3] (total hits) are adjacent
    CK(cudaMemcpyAsync(of, ctr(ctx, 2), sizeof(of), cudaMemcpyDeviceToHost, ctx->stream));
    { int rs = sync_and_check(ctx); if (rs) return rs; }
    if (again) *again = false;
    if (of[0] & 2u) {
        ctx->hits.release();
        CK(ctx->hits.ensure(((size_t)of[2] + 1) * sizeof(uint2)));
        if (again) { *again = true; return DSCGPU_OK; }
        ctx->err = "the z-ray hit buffer was too small for this pass (grown now): repeat the call";
        return DSCGPU_ERR_RAY_OVERFLOW;
    }
    return DSCGPU_OK;
}

// a z-ray pass with its synchronising end; repeated once when the hit buffer had to grow.  h_sums / h_mean (may be null) receive the results.
int run_zray_sync(dscgpu_ctx *ctx, int nb_phase, int mode, const double *c_host, double *d_sums, double *d_mean, double *h_sums, double *h_mean)
{
    for (int attempt = 0; attempt < 2; attempt++) {
        int r = run_zray(ctx, nb_phase, mode, c_host, d_sums, d_mean);
        if (r) return r;
        if (h_sums) CK(cudaMemcpyAsync(h_sums, d_sums, 3 * nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_mean && d_mean) CK(cudaMemcpyAsync(h_mean, d_mean, nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        bool again = false;
        r = check_overflow(ctx, &again);
        if (r) return r;
        if (!again) return DSCGPU_OK;
    }
    ctx->err = "the z-ray hit buffer did not settle";
    return DSCGPU_ERR_RAY_OVERFLOW;
}

int ensure_scratch(dscgpu_ctx *ctx)
{
    CK(ctx->sums.ensure(3 * kMaxPhase * sizeof(double)));
    CK(ctx->partials.ensure((size_t)kMaxPartialRows * 3 * kMaxPhase * sizeof(double)));
    CK(ctx->mean.ensure(kMaxPhase * sizeof(double)));
    CK(ctx->counters.ensure(8 * sizeof(unsigned long long)));
    CK(ctx->misc.ensure(64));
    return DSCGPU_OK;
}

void fill_stats(dscgpu_phase_stats *st, const double *sums, const double *mean, int nb_phase)
{
    memset(st, 0, sizeof(*st));
    for (int k = 0; k < nb_phase; k++) {
        st->total[k] = sums[k];
        st->sumsq[k] = sums[nb_phase + k];
        st->volume[k] = sums[2 * nb_phase + k];
        st->mean[k] = mean[k];
    }
}

// node -> incident interface faces CSR on the device: count, exclusive scan, fill (slot by integer atomic), sort each list
int build_csr_device(dscgpu_ctx *ctx)
{
    const uint32_t nn = ctx->n_nodes, nf = ctx->n_faces;
    CK(ctx->csr_count.ensure(((size_t)nn + 1) * 4));
    CK(ctx->csr_cursor.ensure(((size_t)nn + 1) * 4));
    CK(cudaMemsetAsync(ctx->csr_count.p, 0, ((size_t)nn + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->csr_cursor.p, 0, ((size_t)nn + 1) * 4, ctx->stream));
    const unsigned n_inc = 3u * nf;
    if (n_inc) {
        csr_count_kernel<<<(n_inc + 255) / 256, 256, 0, ctx->stream>>>(ctx->face_nodes.as<unsigned>(), n_inc, nn, ctx->csr_count.as<unsigned>());
        ctx->launches++;
    }
    int rs = exclusive_scan(ctx, ctx->csr_count.as<unsigned>(), ctx->node_off.as<unsigned>(), (size_t)nn + 1, nullptr);
    if (rs) return rs;
    {   // compact list of the nodes that have interface faces (ascending); its length stays on the device (counter 5)
        CK(ctx->csr_flag.ensure(((size_t)nn + 1) * 4));
        CK(ctx->csr_apos.ensure(((size_t)nn + 1) * 4));
        CK(ctx->node_active.ensure(((size_t)nn + 1) * 4));
        csr_flag_kernel<<<(nn + 256) / 256, 256, 0, ctx->stream>>>(ctx->csr_count.as<unsigned>(), nn, ctx->csr_flag.as<unsigned>());
        ctx->launches++;
        rs = exclusive_scan(ctx, ctx->csr_flag.as<unsigned>(), ctx->csr_apos.as<unsigned>(), (size_t)nn + 1, reinterpret_cast<unsigned *>(ctr(ctx, 5)));
        if (rs) return rs;
        if (nn) {
            csr_compact_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->csr_count.as<unsigned>(), ctx->csr_apos.as<unsigned>(), nn,
                                                                         ctx->node_active.as<unsigned>());
            ctx->launches++;
        }
    }
    if (n_inc) {
        csr_fill_kernel<<<(n_inc + 255) / 256, 256, 0, ctx->stream>>>(ctx->face_nodes.as<unsigned>(), n_inc, nn, ctx->node_off.as<unsigned>(),
                                                                     ctx->csr_cursor.as<unsigned>(), ctx->node_inc.as<unsigned>());
        csr_sort_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->node_off.as<unsigned>(), ctx->node_inc.as<unsigned>(), nn);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    // the number of nodes with interface faces, for the compact read-back of the forces (read on the host much later: no wait)
    CK(ctx->h_nact.ensure(sizeof(unsigned)));
    if (!ctx->ev_nact) CK(cudaEventCreateWithFlags(&ctx->ev_nact, cudaEventDisableTiming));
    CK(cudaMemcpyAsync(ctx->h_nact.p, ctr(ctx, 5), sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_nact, ctx->stream));
    return DSCGPU_OK;
}

// Upload of the staged snapshot overlapped with the tet pass: positions first, then the tet arrays in chunks on a copy
// stream; the compute stream validates and integrates chunk i while chunk i+1 is still in flight; faces last (CSR build,
// forces follow).  Leaves the per-phase sums / means on the device like run_tet_pass.
int upload_and_tet_pass_pipelined(dscgpu_ctx *ctx, int nb_phase, int want_sq, double *d_sums_out, double *d_mean_out)
{
    const uint32_t nn = ctx->st_nodes, nt = ctx->st_tets, nf = ctx->st_faces;
    ARG(nf < (1u << 30), "too many interface faces");
    CK(ctx->pos4.ensure((size_t)nn * 32));
    CK(ctx->pos4f.ensure((size_t)nn * 16));
    CK(ctx->tet_nodes.ensure((size_t)nt * 16));
    CK(ctx->tet_label.ensure((size_t)nt * 4));
    CK(ctx->stg_exact.ensure((size_t)nt * 4)); // list of the tets the staged kernel leaves to the exact path
    CK(ctx->face_nodes.ensure((size_t)nf * 12));
    CK(ctx->face_tets.ensure((size_t)nf * 8));
    CK(ctx->face_isb.ensure((size_t)nf));
    CK(ctx->node_off.ensure(((size_t)nn + 1) * 4));
    CK(ctx->node_inc.ensure((size_t)nf * 12));
    CK(ctx->csr_count.ensure(((size_t)nn + 1) * 4));
    CK(ctx->csr_cursor.ensure(((size_t)nn + 1) * 4));
    CK(ctx->h_flag.ensure(sizeof(int)));
    cudaStream_t cs = ctx->copy_stream, ks = ctx->stream;
    int C = ctx->pipeline_chunks;
    if (C > 16) C = 16;
    if (C < 1) C = 1;
    // the copy stream must not overwrite arrays that kernels of the previous evaluation still read
    CK(cudaEventRecord(ctx->ev_chunk[18], ks));
    CK(cudaStreamWaitEvent(cs, ctx->ev_chunk[18], 0));
    if (ctx->timers_on) CK(cudaEventRecord(ctx->ev[DSCGPU_T_H2D][0], cs));
    trace_mark(ctx, "start (copy stream)", cs);
    if (nn) CK(cudaMemcpyAsync(ctx->pos4.p, ctx->h_pos4.p, (size_t)nn * 32, cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(ctx->ev_chunk[16], cs));
    trace_mark(ctx, "positions landed", cs);
    ctx->n_nodes = nn; ctx->n_tets = nt; ctx->n_faces = nf;
    ctx->tet_b = 0; ctx->tet_e = nt; ctx->face_b = 0; ctx->face_e = nf; ctx->node_b = 0; ctx->node_e = nn;
    ctx->partition_stale = false;
    ctx->has_snapshot = true;
    ctx->geom_async = false;
    ctx->init_means_valid = false;
    ctx->forces_resident = ctx->iforces_resident = false;
    ctx->has_face_order = false; // face indices and get_faces(node) belong to one snapshot
    ctx->has_tet_faces = false;
    CK(cudaMemsetAsync(ctr(ctx, 4), 0, sizeof(unsigned long long), ks));
    CK(cudaStreamWaitEvent(ks, ctx->ev_chunk[16], 0));
    if (nn) {
        pos_to_float_kernel<<<(nn + 255) / 256, 256, 0, ks>>>(ctx->pos4.as<double4>(), ctx->pos4f.as<float4>(), nn);
        ctx->launches++;
    }
    int r = tet_pass_begin(ctx);
    if (r) return r;
    TetPassOut none;
    if (ctx->timers_on) CK(cudaEventRecord(ctx->ev[DSCGPU_T_TET][0], ks));
    // Order on the copy stream: tet chunk 0, the (small) face arrays, the other tet chunks.  The compute stream integrates chunk 0 as soon as
    // it is there, builds the node -> face table while chunk 1 is in flight, and works the exact list of every chunk off right behind
    // its staged launch -- so that after the last byte has landed only the last chunk's kernels, the finalize and the forces remain.
    bool faces_done = false;
    auto faces = [&]() -> int {
        if (nf) {
            CK(cudaMemcpyAsync(ctx->face_nodes.p, ctx->h_face_nodes.p, (size_t)nf * 12, cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(ctx->face_tets.p, ctx->h_face_tets.p, (size_t)nf * 8, cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(ctx->face_isb.p, ctx->h_face_isb.p, (size_t)nf, cudaMemcpyHostToDevice, cs));
        }
        CK(cudaEventRecord(ctx->ev_chunk[17], cs));
        trace_mark(ctx, "faces landed", cs);
        faces_done = true;
        return DSCGPU_OK;
    };
    auto faces_device_side = [&]() -> int { // validation of the face ids and the node -> face table, on the compute stream
        CK(cudaStreamWaitEvent(ks, ctx->ev_chunk[17], 0));
        const size_t n_ids = 3 * (size_t)nf + 2 * (size_t)nf;
        if (n_ids) {
            const unsigned grid = (unsigned)std::min<size_t>((n_ids + 255) / 256, (size_t)ctx->num_sms * 16);
            validate_snapshot_kernel<<<grid, 256, 0, ks>>>(nullptr, 0, ctx->face_nodes.as<unsigned>(), 3 * (size_t)nf, ctx->face_tets.as<unsigned>(), 2 * (size_t)nf,
                                                          nn, nt, reinterpret_cast<int *>(ctr(ctx, 4)));
            ctx->launches++;
        }
        const int rc = build_csr_device(ctx);
        trace_mark(ctx, "node->face table built", ks);
        return rc;
    };
    const bool hoist = ctx->pipeline_flags & 1, exact_per_chunk = ctx->pipeline_flags & 2;
    uint32_t chunk_begin[17];
    {
        const double ratio = (ctx->pipeline_flags & 4) ? ctx->pipeline_ratio : 1.0;
        double wsum = 0, w = 1;
        for (int c = 0; c < C; c++) { wsum += w; w *= ratio; }
        double acc = 0;
        w = 1;
        chunk_begin[0] = 0;
        for (int c = 0; c < C; c++) {
            acc += w; w *= ratio;
            uint64_t e = (uint64_t)((double)nt * acc / wsum);
            e = (e + 31) / 32 * 32;
            chunk_begin[c + 1] = c == C - 1 ? nt : (uint32_t)std::min<uint64_t>(e, nt);
        }
    }
    bool first = true;
    for (int c = 0; c < C; c++) {
        // shrinking chunks (multiples of 32 tets): a chunk's kernels finish well inside the copy of the next one, and what is left to do
        // when the last byte has landed is the smallest chunk
        const uint32_t b = chunk_begin[c], e = chunk_begin[c + 1];
        if (e <= b) continue;
        CK(cudaMemcpyAsync(ctx->tet_nodes.as<char>() + (size_t)b * 16, ctx->h_tet_nodes.as<char>() + (size_t)b * 16, (size_t)(e - b) * 16, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(ctx->tet_label.as<char>() + (size_t)b * 4, ctx->h_tet_label.as<char>() + (size_t)b * 4, (size_t)(e - b) * 4, cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(ctx->ev_chunk[c], cs));
        trace_mark(ctx, "tet chunk landed", cs);
        if (first && hoist) { r = faces(); if (r) return r; }
        CK(cudaStreamWaitEvent(ks, ctx->ev_chunk[c], 0));
        const size_t n_ids = 4 * (size_t)(e - b);
        const unsigned grid = (unsigned)std::min<size_t>((n_ids + 255) / 256, (size_t)ctx->num_sms * 16);
        validate_snapshot_kernel<<<grid, 256, 0, ks>>>(ctx->tet_nodes.as<unsigned>() + 4 * (size_t)b, n_ids, nullptr, 0, nullptr, 0, nn, nt,
                                                      reinterpret_cast<int *>(ctr(ctx, 4)));
        ctx->launches++;
        r = tet_pass_range(ctx, nb_phase, 0, 0, nullptr, none, b, e, want_sq);
        if (r) return r;
        trace_mark(ctx, "tet chunk staged done", ks);
        if (exact_per_chunk) { r = launch_exact_list(ctx, /*more_follow=*/true); if (r) return r; }
        if (first && hoist) { r = faces_device_side(); if (r) return r; }
        first = false;
    }
    if (!faces_done) { // (no tets at all, or the face arrays last)
        r = faces(); if (r) return r;
        r = faces_device_side(); if (r) return r;
    }
    if (ctx->timers_on) { CK(cudaEventRecord(ctx->ev[DSCGPU_T_TET][1], ks)); ctx->ev_used[DSCGPU_T_TET] = true; }
    if (ctx->timers_on) { CK(cudaEventRecord(ctx->ev[DSCGPU_T_H2D][1], cs)); ctx->ev_used[DSCGPU_T_H2D] = true; }
    CK(cudaMemcpyAsync(ctx->h_flag.p, ctr(ctx, 4), sizeof(int), cudaMemcpyDeviceToHost, ks));
    ctx->validation_pending = true;
    ctx->staging_valid = false; // uploaded (see commit_snapshot_impl)
    ctx->partition_stale = false;
    r = tet_pass_finish(ctx, nb_phase, d_sums_out, d_mean_out);
    trace_mark(ctx, "tet pass finalised", ks);
    return r;
}

} // namespace

namespace {

CommView comm_view(const dscgpu_ctx *ctx)
{
    CommView c;
    for (int k = 0; k < kCommMaxWorld; k++) c.win[k] = ctx->comm_win[k];
    c.rank = ctx->comm_rank; c.world = ctx->comm_world; c.cap_rows = ctx->comm_cap_rows;
    c.snap_off = comm_snapshot_offset(ctx->comm_world, ctx->comm_cap_rows);
    return c;
}

} // namespace

extern "C" {

int dscgpu_ctx_create(dscgpu_ctx **out, int device, const int dims[3], int dtype, const void *voxels_host)
{
    if (!out) return DSCGPU_ERR_ARG;
    *out = nullptr;
    dscgpu_ctx *ctx = new (std::nothrow) dscgpu_ctx();
    if (!ctx) return DSCGPU_ERR_ARG;
    *out = ctx; // returned even on failure so that dscgpu_last_error can be read; caller destroys it
    for (int i = 0; i < DSCGPU_T_COUNT; i++) { ctx->ev[i][0] = ctx->ev[i][1] = nullptr; ctx->ev_used[i] = false; }
    for (int i = 0; i < 20; i++) ctx->ev_chunk[i] = nullptr;
    ARG(dims && dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "dims must be positive");
    ARG(dtype >= DSCGPU_U8 && dtype <= DSCGPU_F64, "unknown voxel dtype");
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    ARG(device >= 0 && device < n_dev, "no such CUDA device (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    ctx->device = device;
    CK(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 20; i++) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
    if (const char *ev = getenv("DSCGPU_PIPELINE_CHUNKS")) ctx->pipeline_chunks = atoi(ev);
    if (const char *ev = getenv("DSCGPU_PIPELINE_FLAGS")) ctx->pipeline_flags = atoi(ev);
    if (const char *ev = getenv("DSCGPU_E2E_TRACE")) ctx->trace_on = atoi(ev) != 0;
    if (const char *ev = getenv("DSCGPU_PIPELINE_RATIO")) { const double v = atof(ev); if (v > 0.1 && v <= 1.0) ctx->pipeline_ratio = v; }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < DSCGPU_T_COUNT; i++) { CK(cudaEventCreate(&ctx->ev[i][0])); CK(cudaEventCreate(&ctx->ev[i][1])); }
    ctx->dims[0] = dims[0]; ctx->dims[1] = dims[1]; ctx->dims[2] = dims[2];
    ctx->dtype = dtype;
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    CK(cudaMalloc(&ctx->d_vol, n * dtype_size(dtype)));
    if (voxels_host) CK(cudaMemcpyAsync(ctx->d_vol, voxels_host, n * dtype_size(dtype), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMalloc(&ctx->d_tet_bary, sizeof(DSC_TET_BARY)));
    CK(cudaMemcpyAsync(ctx->d_tet_bary, DSC_TET_BARY, sizeof(DSC_TET_BARY), cudaMemcpyHostToDevice, ctx->stream));
    {
        std::vector<float> tf(4 * DSC_TET_POINTS_TOTAL);
        for (int i = 0; i < 4 * DSC_TET_POINTS_TOTAL; i++) tf[i] = (float)DSC_TET_BARY[i];
        CK(cudaMalloc(&ctx->d_tet_bary_f, tf.size() * sizeof(float)));
        CK(cudaMemcpy(ctx->d_tet_bary_f, tf.data(), tf.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    {
        // lattice form of the sample table: b_k = a_k/(4n) + eta_k, integer a_k summing to 4n, |eta_k| <= 5e-7, eps = sum b_k - 1
        std::vector<float> tl(4 * DSC_TET_POINTS_TOTAL);
        for (int L = 0; L < DSC_TET_LEVELS; L++) {
            const int n = L + 1;
            for (int i = DSC_TET_LEVEL_OFFSET[L]; i < DSC_TET_LEVEL_OFFSET[L + 1]; i++) {
                const double *b = DSC_TET_BARY + 4 * i;
                long asum = 0;
                double av[4];
                for (int k = 0; k < 4; k++) {
                    av[k] = (double)lrint(4.0 * n * b[k]);
                    asum += (long)av[k];
                    ARG(fabs(b[k] - av[k] / (4.0 * n)) <= 5.0e-7, "sample table is not the expected 6-decimal lattice");
                    ARG(av[k] >= 1.0, "sample table row on the boundary of the tet"); // w2's in-volume argument needs every weight >= 1/(4n) - 5e-7
                }
                ARG(asum == 4 * n, "sample table row does not sum to one on the lattice");
                const double eps = ((b[0] + b[1]) + b[2]) + b[3] - 1.0;
                ARG(fabs(eps) <= 2.1e-6, "sample table row sum too far from one");
                tl[4 * i] = (float)av[1]; tl[4 * i + 1] = (float)av[2]; tl[4 * i + 2] = (float)av[3]; tl[4 * i + 3] = (float)eps;
            }
        }
        CK(cudaMalloc(&ctx->d_tet_lat, tl.size() * sizeof(float)));
        CK(cudaMemcpy(ctx->d_tet_lat, tl.data(), tl.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (const char *ev = getenv("DSCGPU_TET_VARIANT")) ctx->tet_variant = atoi(ev);
    // staged kernel: voxel box of one shared-memory slot by voxel size (a 32-tet tile of a DSC mesh whose tets hold 27-64 voxels: about
    // 52 x 11 x 11 voxels), as many slots as fit -- the more slots, the more TMA transfers in flight next to the consumers' tiles
    ctx->stg_slot_bytes = (ctx->dtype == DSCGPU_U8 ? 8 : ctx->dtype == DSCGPU_U16 ? 14 : 24) * 1024;
    ctx->stg_slots = kStgMaxSlots;
    if (const char *ev = getenv("DSCGPU_STG_SLOT_KB")) { // development knobs
        const int kb = atoi(ev);
        if (kb >= 4 && kb <= 64) ctx->stg_slot_bytes = kb * 1024;
    }
    if (const char *ev = getenv("DSCGPU_STG_SLOTS")) {
        const int n = atoi(ev);
        if (n >= 2 && n <= kStgMaxSlots) ctx->stg_slots = n;
    }
    while (ctx->stg_slots > 2 && ctx->stg_slots * (ctx->stg_slot_bytes + kStgRecBytes) > 227 * 1024 - 4096) ctx->stg_slots--;
    if (voxels_host) { // one scan of a float volume: are all voxels non-negative finite numbers?
        int rf = ensure_volume_flags(ctx);
        if (rf) return rf;
    }
    if (const char *ev = getenv("DSCGPU_FORCE_THREAD_GATHER")) ctx->force_thread_gather = atoi(ev);
    if (const char *ev = getenv("DSCGPU_FORCE_NODE_FORM")) ctx->force_node_form = atoi(ev) == 2 ? 2 : 3;
    CK(cudaMemcpyToSymbolAsync(c_gauss, DSC_GAUSS, sizeof(DSC_GAUSS), 0, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyToSymbolAsync(c_gauss_off, DSC_GAUSS_SET_OFFSET, sizeof(DSC_GAUSS_SET_OFFSET), 0, cudaMemcpyHostToDevice, ctx->stream));
    int r = ensure_scratch(ctx);
    if (r) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return DSCGPU_OK;
}

int dscgpu_ctx_destroy(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_OK;
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    DevBuf *bufs[] = {&ctx->fgeo, &ctx->flab, &ctx->fint, &ctx->upd, &ctx->nflags, &ctx->iterm, &ctx->inorm, &ctx->iskip, &ctx->iforce, &ctx->tet_face_nodes, &ctx->tet_face_tets, &ctx->csr_flag, &ctx->csr_apos, &ctx->node_active, &ctx->csr_count, &ctx->csr_cursor, &ctx->fbuf, &ctx->fset, &ctx->pos4, &ctx->pos4f, &ctx->tet_nodes, &ctx->tet_label, &ctx->face_nodes, &ctx->face_tets, &ctx->face_isb, &ctx->node_off, &ctx->node_inc,
                      &ctx->partials, &ctx->sums, &ctx->mean, &ctx->forces, &ctx->counters, &ctx->o_total, &ctx->o_vol, &ctx->o_mean, &ctx->o_energy,
                      &ctx->o_variation, &ctx->o_hash, &ctx->o_nsamp, &ctx->ray_count, &ctx->ray_offset, &ctx->ray_cursor, &ctx->hits, &ctx->block_sums,
                      &ctx->ray_psum, &ctx->ray_pcnt, &ctx->misc, &ctx->vbound, &ctx->pit, &ctx->stg_exact, &ctx->ndisp, &ctx->ndest, &ctx->crows};
    for (DevBuf *b : bufs) b->release();
    PinBuf *pins[] = {&ctx->h_pos4, &ctx->h_tet_nodes, &ctx->h_tet_label, &ctx->h_face_nodes, &ctx->h_face_tets, &ctx->h_face_isb, &ctx->h_node_off,
                      &ctx->h_node_inc, &ctx->h_node_active, &ctx->h_out, &ctx->h_flag, &ctx->h_upd, &ctx->h_nact};
    for (PinBuf *b : pins) b->release();
    if (ctx->d_vol) cudaFree(ctx->d_vol);
    if (ctx->d_sum_table) cudaFree(ctx->d_sum_table);
    if (ctx->d_sq_table) cudaFree(ctx->d_sq_table);
    if (ctx->d_tet_bary) cudaFree(ctx->d_tet_bary);
    if (ctx->d_tet_bary_f) cudaFree(ctx->d_tet_bary_f);
    if (ctx->d_tet_lat) cudaFree(ctx->d_tet_lat);
    for (int k = 0; k < 8; k++) {
        if (!ctx->comm_win[k]) continue;
        if (ctx->comm_ipc[k]) cudaIpcCloseMemHandle(ctx->comm_win[k]);
        else if (k == ctx->comm_rank && ctx->comm_world) cudaFree(ctx->comm_win[k]);
    }
    if (ctx->comm_rows) cudaFree(ctx->comm_rows);
    if (ctx->d_vol_flags) cudaFree(ctx->d_vol_flags);
    if (ctx->d_stg_maps) cudaFree(ctx->d_stg_maps);
    if (ctx->d_stg_counter) cudaFree(ctx->d_stg_counter);
    for (int i = 0; i < DSCGPU_T_COUNT; i++) {
        if (ctx->ev[i][0]) cudaEventDestroy(ctx->ev[i][0]);
        if (ctx->ev[i][1]) cudaEventDestroy(ctx->ev[i][1]);
    }
    for (int i = 0; i < 20; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    for (cudaEvent_t e : ctx->trace_pool) cudaEventDestroy(e);
    if (ctx->ev_nact) cudaEventDestroy(ctx->ev_nact);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return DSCGPU_OK;
}

const char *dscgpu_last_error(const dscgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dscgpu_set_stream(dscgpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return DSCGPU_OK;
}

int dscgpu_synchronize(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    return sync_and_check(ctx);
}

void *dscgpu_volume_ptr_dev(dscgpu_ctx *ctx)
{
    if (!ctx) return nullptr;
    ctx->vol_flags_valid = false;
    ctx->has_sq_table = false;
    ctx->stg_state = 0; // the fixed-point scales depend on the largest voxel
    return ctx->d_vol;
}

int dscgpu_volume_modified(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->vol_flags_valid = false;
    ctx->has_sq_table = false;
    ctx->stg_state = 0; // the fixed-point scales depend on the largest voxel
    return DSCGPU_OK;
}

int dscgpu_volume_upload(dscgpu_ctx *ctx, const void *voxels_host)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(voxels_host, "voxels_host is NULL");
    const size_t n = (size_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
    CK(cudaMemcpyAsync(ctx->d_vol, voxels_host, n * dtype_size(ctx->dtype), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->vol_flags_valid = false;
    ctx->stg_state = 0; // the fixed-point scales depend on the largest voxel
    ctx->has_sum_table = false; ctx->has_sq_table = false;
    return DSCGPU_OK;
}
double *dscgpu_sum_table_ptr_dev(dscgpu_ctx *ctx) { return ctx ? ctx->d_sum_table : nullptr; }

int dscgpu_build_sum_table(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    size_t n = (size_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
    if (!ctx->d_sum_table) CK(cudaMalloc(&ctx->d_sum_table, n * sizeof(double)));
    size_t XY = (size_t)ctx->dims[0] * ctx->dims[1];
    unsigned grid = (unsigned)((XY + 127) / 128);
    TimerScope ts(ctx, DSCGPU_T_SUMTABLE);
    switch (ctx->dtype) {
    case DSCGPU_U8: sum_table_kernel<unsigned char><<<grid, 128, 0, ctx->stream>>>((const unsigned char *)ctx->d_vol, ctx->d_sum_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
    case DSCGPU_U16: sum_table_kernel<unsigned short><<<grid, 128, 0, ctx->stream>>>((const unsigned short *)ctx->d_vol, ctx->d_sum_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
    case DSCGPU_F32: sum_table_kernel<float><<<grid, 128, 0, ctx->stream>>>((const float *)ctx->d_vol, ctx->d_sum_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
    default: sum_table_kernel<double><<<grid, 128, 0, ctx->stream>>>((const double *)ctx->d_vol, ctx->d_sum_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->has_sum_table = true;
    ctx->has_sq_table = false; // built from the same voxels, on the next energy read-out
    return DSCGPU_OK;
}

int dscgpu_get_sum_table(dscgpu_ctx *ctx, double *table_host)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    if (!ctx->has_sum_table) { ctx->err = "sum table not built"; return DSCGPU_ERR_NO_SUMTABLE; }
    size_t n = (size_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
    CK(cudaMemcpyAsync(table_host, ctx->d_sum_table, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSCGPU_OK;
}

int dscgpu_snapshot_staging_get(dscgpu_ctx *ctx, uint32_t n_nodes, uint32_t n_tets, uint32_t n_faces, dscgpu_snapshot_staging *out)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(out, "null staging struct");
    // the previous upload may still be reading the pinned arrays
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->h_pos4.ensure((size_t)n_nodes * 4 * sizeof(double)));
    CK(ctx->h_tet_nodes.ensure((size_t)n_tets * 4 * sizeof(uint32_t)));
    CK(ctx->h_tet_label.ensure((size_t)n_tets * sizeof(int32_t)));
    CK(ctx->h_face_nodes.ensure((size_t)n_faces * 3 * sizeof(uint32_t)));
    CK(ctx->h_face_tets.ensure((size_t)n_faces * 2 * sizeof(uint32_t)));
    CK(ctx->h_face_isb.ensure((size_t)n_faces));
    ctx->st_nodes = n_nodes; ctx->st_tets = n_tets; ctx->st_faces = n_faces;
    ctx->staging_valid = true;
    out->pos4 = ctx->h_pos4.as<double>();
    out->tet_nodes = ctx->h_tet_nodes.as<uint32_t>();
    out->tet_label = ctx->h_tet_label.as<int32_t>();
    out->face_nodes = ctx->h_face_nodes.as<uint32_t>();
    out->face_tets = ctx->h_face_tets.as<uint32_t>();
    out->face_is_boundary = ctx->h_face_isb.as<uint8_t>();
    return DSCGPU_OK;
}

static int commit_snapshot_impl(dscgpu_ctx *ctx, bool sharded);

int dscgpu_commit_snapshot(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    return commit_snapshot_impl(ctx, false);
}

int dscgpu_commit_snapshot_sharded(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(ctx->comm_connected && ctx->comm_snap_bytes, "dscgpu_comm_create_ex with a snapshot region, and dscgpu_comm_connect, first");
    return commit_snapshot_impl(ctx, true);
}

static int commit_snapshot_impl(dscgpu_ctx *ctx, bool sharded)
{
    ARG(ctx->staging_valid, "dscgpu_snapshot_staging_get was not called");
    const uint32_t nn = ctx->st_nodes, nt = ctx->st_tets, nf = ctx->st_faces;
    ARG(nf < (1u << 30), "too many interface faces");
    const uint32_t *tn = ctx->h_tet_nodes.as<uint32_t>();
    const uint32_t *fn = ctx->h_face_nodes.as<uint32_t>();
    CK(ctx->pos4.ensure((size_t)nn * 32));
    CK(ctx->pos4f.ensure((size_t)nn * 16));
    CK(ctx->tet_nodes.ensure((size_t)nt * 16));
    CK(ctx->tet_label.ensure((size_t)nt * 4));
    CK(ctx->stg_exact.ensure((size_t)nt * 4)); // list of the tets the staged kernel leaves to the exact path
    CK(ctx->face_nodes.ensure((size_t)nf * 12));
    CK(ctx->face_tets.ensure((size_t)nf * 8));
    CK(ctx->face_isb.ensure((size_t)nf));
    CK(ctx->node_off.ensure(((size_t)nn + 1) * 4));
    CK(ctx->node_inc.ensure((size_t)nf * 12));
    if (sharded) {
        // rank r sends bytes [T r / W, T (r+1) / W) of the flattened snapshot (six arrays at 256-byte aligned offsets) over PCIe into its
        // window; the slices travel to the peers over NVLink (dsc_comm.cuh); the assembled region is copied into the snapshot buffers
        TimerScope ts(ctx, DSCGPU_T_H2D);
        cudaStream_t st = ctx->stream;
        const size_t len[6] = {(size_t)nn * 32, (size_t)nt * 16, (size_t)nt * 4, (size_t)nf * 12, (size_t)nf * 8, (size_t)nf};
        const void *hsrc[6] = {ctx->h_pos4.p, ctx->h_tet_nodes.p, ctx->h_tet_label.p, ctx->h_face_nodes.p, ctx->h_face_tets.p, ctx->h_face_isb.p};
        void *ddst[6] = {ctx->pos4.p, ctx->tet_nodes.p, ctx->tet_label.p, ctx->face_nodes.p, ctx->face_tets.p, ctx->face_isb.p};
        size_t off[7];
        off[0] = 0;
        for (int k = 0; k < 6; k++) off[k + 1] = (off[k] + len[k] + 255) & ~(size_t)255;
        const size_t T = off[6];
        ARG(T <= ctx->comm_snap_bytes, "the snapshot region of the exchange window is too small for this snapshot (dscgpu_comm_create_ex)");
        const int W = ctx->comm_world, R = ctx->comm_rank;
        const size_t b = (T / 256 * R / W) * 256, e = R == W - 1 ? T : (T / 256 * (R + 1) / W) * 256;
        unsigned char *region = ctx->comm_win[R] + comm_snapshot_offset(W, ctx->comm_cap_rows);
        for (int k = 0; k < 6; k++) { // the part of array k inside [b, e)
            const size_t lo = std::max(b, off[k]), hi = std::min(e, off[k] + len[k]);
            if (hi > lo) CK(cudaMemcpyAsync(region + lo, static_cast<const unsigned char *>(hsrc[k]) + (lo - off[k]), hi - lo, cudaMemcpyHostToDevice, st));
        }
        if (W > 1) {
            unsigned grid = (unsigned)std::min<size_t>(((e - b) / 16 + 255) / 256 + 1, (size_t)ctx->num_sms * 2);
            comm_snapshot_push_kernel<<<grid, 256, 0, st>>>(comm_view(ctx), b, e);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        SnapshotParts parts;
        for (int k = 0; k < 6; k++) { parts.dst[k] = static_cast<unsigned char *>(ddst[k]); parts.off[k] = off[k]; parts.len[k] = len[k]; }
        comm_snapshot_gather_kernel<<<(unsigned)std::min<size_t>((T / 16 + 255) / 256 + 1, (size_t)ctx->num_sms * 4), 256, 0, st>>>(comm_view(ctx), parts);
        ctx->launches++;
        CK(cudaGetLastError());
    } else {
        TimerScope ts(ctx, DSCGPU_T_H2D);
        cudaStream_t s = ctx->stream;
        if (nn) CK(cudaMemcpyAsync(ctx->pos4.p, ctx->h_pos4.p, (size_t)nn * 32, cudaMemcpyHostToDevice, s));
        if (nt) CK(cudaMemcpyAsync(ctx->tet_nodes.p, ctx->h_tet_nodes.p, (size_t)nt * 16, cudaMemcpyHostToDevice, s));
        if (nt) CK(cudaMemcpyAsync(ctx->tet_label.p, ctx->h_tet_label.p, (size_t)nt * 4, cudaMemcpyHostToDevice, s));
        if (nf) CK(cudaMemcpyAsync(ctx->face_nodes.p, ctx->h_face_nodes.p, (size_t)nf * 12, cudaMemcpyHostToDevice, s));
        if (nf) CK(cudaMemcpyAsync(ctx->face_tets.p, ctx->h_face_tets.p, (size_t)nf * 8, cudaMemcpyHostToDevice, s));
        if (nf) CK(cudaMemcpyAsync(ctx->face_isb.p, ctx->h_face_isb.p, (size_t)nf, cudaMemcpyHostToDevice, s));
    }
    ctx->n_nodes = nn; ctx->n_tets = nt; ctx->n_faces = nf;
    {
        int rs = build_csr_device(ctx);
        if (rs) return rs;
    }
    if (nn) {
        pos_to_float_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->pos4.as<double4>(), ctx->pos4f.as<float4>(), nn);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->n_nodes = nn; ctx->n_tets = nt; ctx->n_faces = nf;
    // index validation runs on the device (a host pass over 8 M ids would dominate the end-to-end step): bad ids are
    // replaced by 0 / NO_TET so that no kernel can read out of range, and the error is raised at the next synchronisation
    {
        CK(ctx->h_flag.ensure(sizeof(int)));
        CK(cudaMemsetAsync(ctr(ctx, 4), 0, sizeof(unsigned long long), ctx->stream));
        size_t n_ids = 4 * (size_t)nt + 3 * (size_t)nf + 2 * (size_t)nf;
        if (n_ids) {
            unsigned grid = (unsigned)std::min<size_t>((n_ids + 255) / 256, (size_t)ctx->num_sms * 16);
            validate_snapshot_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->tet_nodes.as<unsigned>(), 4 * (size_t)nt, ctx->face_nodes.as<unsigned>(),
                                                                  3 * (size_t)nf, ctx->face_tets.as<unsigned>(), 2 * (size_t)nf, nn, nt,
                                                                  reinterpret_cast<int *>(ctr(ctx, 4)));
            ctx->launches++;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(ctx->h_flag.p, ctr(ctx, 4), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->validation_pending = true;
    }
    ctx->tet_b = 0; ctx->tet_e = nt; ctx->face_b = 0; ctx->face_e = nf; ctx->node_b = 0; ctx->node_e = nn;
    ctx->partition_stale = false;
    ctx->has_snapshot = true;
    ctx->geom_async = false;
    ctx->init_means_valid = false;
    ctx->forces_resident = ctx->iforces_resident = false;
    ctx->has_face_order = false; // face indices and get_faces(node) belong to one snapshot
    ctx->has_tet_faces = false;

    // choose lanes per tet from a sample of tet volumes (heuristic only; never affects results)
    {
        const double *p4 = ctx->h_pos4.as<double>();
        double acc = 0;
        int cnt = 0;
        uint32_t step = nt / 1024 + 1;
        for (uint32_t t = 0; t < nt; t += step) {
            const uint32_t *q = tn + 4 * (size_t)t;
            if (q[0] >= nn || q[1] >= nn || q[2] >= nn || q[3] >= nn) continue;
            V3 a{p4[4 * (size_t)q[0]], p4[4 * (size_t)q[0] + 1], p4[4 * (size_t)q[0] + 2]};
            V3 b{p4[4 * (size_t)q[1]], p4[4 * (size_t)q[1] + 1], p4[4 * (size_t)q[1] + 2]};
            V3 c{p4[4 * (size_t)q[2]], p4[4 * (size_t)q[2] + 1], p4[4 * (size_t)q[2] + 2]};
            V3 d{p4[4 * (size_t)q[3]], p4[4 * (size_t)q[3] + 1], p4[4 * (size_t)q[3] + 2]};
            int L = tet_level(tet_volume(a, b, c, d));
            acc += (L + 1) * (L + 1) * (L + 1);
            cnt++;
        }
        ctx->auto_lanes = (cnt && acc / cnt >= 200.0) ? 32 : 8;
    }
    ctx->staging_valid = false; // uploaded: dscgpu_image_force_eval(ctx, NULL, ...) evaluates the resident snapshot until the staging is taken again
    return DSCGPU_OK;
}

// layout of the pinned update block: [keys | pos | tet idx | tet nodes | tet labels | face nodes | face tets | face flags]
namespace {
struct UpdLayout {
    size_t keys, pos, tidx, tn, tl, fn, ft, fb, total;
};
UpdLayout upd_layout(uint32_t n_pos, bool with_keys, uint32_t n_tet, uint32_t n_faces)
{
    auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
    UpdLayout L;
    L.keys = 0;
    L.pos = up(with_keys ? (size_t)n_pos * 4 : 0);
    L.tidx = up(L.pos + (size_t)n_pos * 24);
    L.tn = up(L.tidx + (size_t)n_tet * 4);
    L.tl = up(L.tn + (size_t)n_tet * 16);
    L.fn = up(L.tl + (size_t)n_tet * 4);
    L.ft = up(L.fn + (size_t)n_faces * 12);
    L.fb = up(L.ft + (size_t)n_faces * 8);
    L.total = up(L.fb + (size_t)n_faces) + 16;
    return L;
}
} // namespace

int dscgpu_update_staging_get(dscgpu_ctx *ctx, uint32_t n_pos, int with_keys, uint32_t n_tet_upd, uint32_t n_faces, dscgpu_update_staging *out)
{
    NEED_SNAPSHOT();
    ARG(out, "out is null");
    ARG(n_faces < (1u << 30), "too many interface faces");
    const UpdLayout L = upd_layout(n_pos, with_keys != 0, n_tet_upd, n_faces);
    CK(cudaStreamSynchronize(ctx->stream)); // the block may still feed the previous update
    CK(ctx->h_upd.ensure(L.total));
    unsigned char *h = ctx->h_upd.as<unsigned char>();
    out->node_keys = with_keys ? reinterpret_cast<uint32_t *>(h + L.keys) : nullptr;
    out->pos = reinterpret_cast<double *>(h + L.pos);
    out->tet_index = reinterpret_cast<uint32_t *>(h + L.tidx);
    out->tet_nodes = reinterpret_cast<uint32_t *>(h + L.tn);
    out->tet_label = reinterpret_cast<int32_t *>(h + L.tl);
    out->face_nodes = reinterpret_cast<uint32_t *>(h + L.fn);
    out->face_tets = reinterpret_cast<uint32_t *>(h + L.ft);
    out->face_is_boundary = h + L.fb;
    ctx->upd_n_pos = n_pos; ctx->upd_keys = with_keys != 0; ctx->upd_n_tet = n_tet_upd; ctx->upd_n_faces = n_faces;
    ctx->upd_staged = true;
    return DSCGPU_OK;
}

int dscgpu_update_commit(dscgpu_ctx *ctx, int replace_faces)
{
    NEED_SNAPSHOT();
    ARG(ctx->upd_staged, "dscgpu_update_staging_get was not called");
    const uint32_t nn = ctx->n_nodes, nt = ctx->n_tets;
    const uint32_t n_pos = ctx->upd_n_pos, n_tet_upd = ctx->upd_n_tet, n_faces = replace_faces ? ctx->upd_n_faces : 0;
    const UpdLayout L = upd_layout(n_pos, ctx->upd_keys, n_tet_upd, ctx->upd_n_faces);
    CK(ctx->upd.ensure(L.total));
    CK(ctx->h_flag.ensure(sizeof(int)));
    ctx->forces_resident = ctx->iforces_resident = false; // the mesh moves: forces of the previous state do not apply
    if (replace_faces) ctx->has_face_order = false;       // indices into the face list that is being replaced
    TimerScope ts(ctx, DSCGPU_T_H2D);
    // one H2D copy for everything that is scattered; the face arrays go straight to their place
    const size_t scatter_bytes = L.fn;
    if (n_pos || n_tet_upd) CK(cudaMemcpyAsync(ctx->upd.p, ctx->h_upd.p, scatter_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->validation_pending) CK(cudaMemsetAsync(ctr(ctx, 4), 0, sizeof(unsigned long long), ctx->stream));
    unsigned char *d = ctx->upd.as<unsigned char>();
    const unsigned char *h = ctx->h_upd.as<unsigned char>();
    if (n_pos) {
        scatter_positions_kernel<<<(n_pos + 255) / 256, 256, 0, ctx->stream>>>(ctx->upd_keys ? reinterpret_cast<const unsigned *>(d + L.keys) : nullptr,
                                                                              reinterpret_cast<const double *>(d + L.pos), n_pos, ctx->pos4.as<double4>(),
                                                                              ctx->pos4f.as<float4>(), nn, reinterpret_cast<int *>(ctr(ctx, 4)));
        ctx->launches++;
    }
    if (n_tet_upd) {
        scatter_tets_kernel<<<(n_tet_upd + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned *>(d + L.tidx), reinterpret_cast<const uint4 *>(d + L.tn),
                                                                             reinterpret_cast<const int *>(d + L.tl), n_tet_upd, ctx->tet_nodes.as<uint4>(),
                                                                             ctx->tet_label.as<int>(), nt, nn, reinterpret_cast<int *>(ctr(ctx, 4)));
        ctx->launches++;
        ctx->has_tet_faces = false;
    }
    CK(cudaGetLastError());
    if (replace_faces) {
        const uint32_t nf = n_faces;
        CK(ctx->face_nodes.ensure((size_t)nf * 12));
        CK(ctx->face_tets.ensure((size_t)nf * 8));
        CK(ctx->face_isb.ensure((size_t)nf));
        CK(ctx->node_inc.ensure((size_t)nf * 12));
        if (nf) {
            CK(cudaMemcpyAsync(ctx->face_nodes.p, h + L.fn, (size_t)nf * 12, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->face_tets.p, h + L.ft, (size_t)nf * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->face_isb.p, h + L.fb, (size_t)nf, cudaMemcpyHostToDevice, ctx->stream));
        }
        // the partition is one unit (tets, faces, nodes): the full ranges follow the new face count; under a real partition an unchanged
        // face count keeps the ranges, a changed one makes them stale -- evaluating all faces on every rank would add the forces world times
        const bool whole = ctx->tet_b == 0 && ctx->tet_e == ctx->n_tets && ctx->node_b == 0 && ctx->node_e == ctx->n_nodes && ctx->face_b == 0 && ctx->face_e == ctx->n_faces;
        if (whole) { ctx->face_b = 0; ctx->face_e = nf; }
        else if (nf != ctx->n_faces) { ctx->partition_stale = true; ctx->face_b = ctx->face_e = 0; }
        ctx->n_faces = nf;
        int rs = build_csr_device(ctx);
        if (rs) return rs;
        if (nf) {
            const size_t n_ids = 5 * (size_t)nf;
            unsigned grid = (unsigned)std::min<size_t>((n_ids + 255) / 256, (size_t)ctx->num_sms * 16);
            validate_snapshot_kernel<<<grid, 256, 0, ctx->stream>>>(nullptr, 0, ctx->face_nodes.as<unsigned>(), 3 * (size_t)nf, ctx->face_tets.as<unsigned>(),
                                                                  2 * (size_t)nf, nn, nt, reinterpret_cast<int *>(ctr(ctx, 4)));
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    CK(cudaMemcpyAsync(ctx->h_flag.p, ctr(ctx, 4), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->validation_pending = true;
    ctx->init_means_valid = false;
    ctx->geom_async = false;
    ctx->staging_valid = false; // the pinned staging of the last full snapshot is stale: dscgpu_image_force_eval(ctx, NULL, ...) must not re-upload it
    return DSCGPU_OK;
}

int dscgpu_update_snapshot(dscgpu_ctx *ctx, uint32_t n_pos, const uint32_t *node_keys, const double *pos, uint32_t n_tet_upd, const uint32_t *tet_index,
                           const uint32_t *tet_nodes, const int32_t *tet_label, int replace_faces, uint32_t n_faces, const uint32_t *face_nodes,
                           const uint32_t *face_tets, const uint8_t *face_is_boundary)
{
    NEED_SNAPSHOT();
    ARG(n_pos == 0 || pos, "pos is null");
    ARG(n_tet_upd == 0 || (tet_index && tet_nodes && tet_label), "tet update arrays are null");
    ARG(!replace_faces || n_faces == 0 || (face_nodes && face_tets && face_is_boundary), "face arrays are null");
    dscgpu_update_staging st;
    int r = dscgpu_update_staging_get(ctx, n_pos, node_keys != nullptr, n_tet_upd, replace_faces ? n_faces : 0, &st);
    if (r) return r;
    if (node_keys && n_pos) memcpy(st.node_keys, node_keys, (size_t)n_pos * 4);
    if (n_pos) memcpy(st.pos, pos, (size_t)n_pos * 24);
    if (n_tet_upd) {
        memcpy(st.tet_index, tet_index, (size_t)n_tet_upd * 4);
        memcpy(st.tet_nodes, tet_nodes, (size_t)n_tet_upd * 16);
        memcpy(st.tet_label, tet_label, (size_t)n_tet_upd * 4);
    }
    if (replace_faces && n_faces) {
        memcpy(st.face_nodes, face_nodes, (size_t)n_faces * 12);
        memcpy(st.face_tets, face_tets, (size_t)n_faces * 8);
        memcpy(st.face_is_boundary, face_is_boundary, (size_t)n_faces);
    }
    return dscgpu_update_commit(ctx, replace_faces);
}

int dscgpu_set_snapshot(dscgpu_ctx *ctx, const dscgpu_snapshot *s)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(s, "null snapshot");
    ARG(s->n_nodes_buffer == 0 || s->pos, "pos is null");
    ARG(s->n_tets == 0 || (s->tet_nodes && s->tet_label), "tet arrays are null");
    ARG(s->n_faces == 0 || (s->face_nodes && s->face_tets && s->face_is_boundary), "face arrays are null");
    dscgpu_snapshot_staging st;
    int r = dscgpu_snapshot_staging_get(ctx, s->n_nodes_buffer, s->n_tets, s->n_faces, &st);
    if (r) return r;
    for (size_t i = 0; i < s->n_nodes_buffer; i++) {
        st.pos4[4 * i] = s->pos[3 * i]; st.pos4[4 * i + 1] = s->pos[3 * i + 1]; st.pos4[4 * i + 2] = s->pos[3 * i + 2]; st.pos4[4 * i + 3] = 0.0;
    }
    if (s->n_tets) {
        memcpy(st.tet_nodes, s->tet_nodes, (size_t)s->n_tets * 16);
        memcpy(st.tet_label, s->tet_label, (size_t)s->n_tets * 4);
    }
    if (s->n_faces) {
        memcpy(st.face_nodes, s->face_nodes, (size_t)s->n_faces * 12);
        memcpy(st.face_tets, s->face_tets, (size_t)s->n_faces * 8);
        memcpy(st.face_is_boundary, s->face_is_boundary, (size_t)s->n_faces);
    }
    r = dscgpu_commit_snapshot(ctx);
    if (r) return r;
    return sync_and_check(ctx); // this entry point reports a bad snapshot immediately
}

int dscgpu_set_partition(dscgpu_ctx *ctx, uint32_t tb, uint32_t te, uint32_t fb, uint32_t fe, uint32_t nb, uint32_t ne)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    if (!ctx->has_snapshot) { ctx->err = "no snapshot"; return DSCGPU_ERR_NO_SNAPSHOT; }
    ARG(tb <= te && te <= ctx->n_tets && fb <= fe && fe <= ctx->n_faces && nb <= ne && ne <= ctx->n_nodes, "partition out of range");
    ctx->tet_b = tb; ctx->tet_e = te; ctx->face_b = fb; ctx->face_e = fe; ctx->node_b = nb; ctx->node_e = ne;
    ctx->partition_stale = false;
    return DSCGPU_OK;
}


int dscgpu_tet_integrate(dscgpu_ctx *ctx, int nb_phase, int include_bound, int nc, const double *c, double *tet_total, double *tet_vol,
                         double *tet_mean, double *tet_energy, double *tet_variation, dscgpu_phase_stats *stats)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE, "nb_phase out of range");
    ARG(nc >= 0 && nc <= DSCGPU_MAX_PHASE && (nc == 0 || c), "nc out of range or c null");
    const size_t nt = ctx->n_tets;
    double *d_total = nullptr, *d_vol = nullptr, *d_mean = nullptr, *d_en = nullptr, *d_var = nullptr;
    if (tet_total) { CK(ctx->o_total.ensure(nt * 8)); d_total = ctx->o_total.as<double>(); }
    if (tet_vol) { CK(ctx->o_vol.ensure(nt * 8)); d_vol = ctx->o_vol.as<double>(); }
    if (tet_mean) { CK(ctx->o_mean.ensure(nt * 8)); d_mean = ctx->o_mean.as<double>(); }
    if (tet_energy && nc) { CK(ctx->o_energy.ensure(nt * nc * 8)); d_en = ctx->o_energy.as<double>(); }
    if (tet_variation && nc) { CK(ctx->o_variation.ensure(nt * nc * 8)); d_var = ctx->o_variation.as<double>(); }
    // tets outside the partition keep -1
    if (d_total && !(ctx->tet_b == 0 && ctx->tet_e == ctx->n_tets)) {
        // (partitioned per-tet outputs are only used by tests; fill with a recognisable pattern)
        CK(cudaMemsetAsync(d_total, 0xFF, nt * 8, ctx->stream));
    }
    int r = run_tet_pass(ctx, nb_phase, include_bound, nc, c, d_total, d_vol, d_mean, d_en, d_var, ctx->sums.as<double>(), ctx->mean.as<double>());
    if (r) return r;
    {
        TimerScope ts(ctx, DSCGPU_T_D2H);
        cudaStream_t s = ctx->stream;
        if (d_total) CK(cudaMemcpyAsync(tet_total, d_total, nt * 8, cudaMemcpyDeviceToHost, s));
        if (d_vol) CK(cudaMemcpyAsync(tet_vol, d_vol, nt * 8, cudaMemcpyDeviceToHost, s));
        if (d_mean) CK(cudaMemcpyAsync(tet_mean, d_mean, nt * 8, cudaMemcpyDeviceToHost, s));
        if (d_en) CK(cudaMemcpyAsync(tet_energy, d_en, nt * nc * 8, cudaMemcpyDeviceToHost, s));
        if (d_var) CK(cudaMemcpyAsync(tet_variation, d_var, nt * nc * 8, cudaMemcpyDeviceToHost, s));
    }
    double h_sums[3 * kMaxPhase], h_mean[kMaxPhase];
    CK(cudaMemcpyAsync(h_sums, ctx->sums.p, 3 * nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_mean, ctx->mean.p, nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    { int rs = sync_and_check(ctx); if (rs) return rs; }
    if (stats) fill_stats(stats, h_sums, h_mean, nb_phase);
    return DSCGPU_OK;
}

int dscgpu_set_tet_faces(dscgpu_ctx *ctx, const uint32_t *tet_face_nodes, const uint32_t *tet_face_tets)
{
    NEED_SNAPSHOT();
    ARG(tet_face_nodes && tet_face_tets, "null argument");
    const size_t nt = ctx->n_tets;
    for (size_t i = 0; i < 12 * nt; i++) ARG(tet_face_nodes[i] < ctx->n_nodes, "tet_face_nodes holds a node key >= n_nodes_buffer");
    for (size_t i = 0; i < 8 * nt; i++) ARG(tet_face_tets[i] < ctx->n_tets || tet_face_tets[i] == DSCGPU_NO_TET, "tet_face_tets holds a tet index >= n_tets");
    CK(ctx->tet_face_nodes.ensure(nt * 48 + 4));
    CK(ctx->tet_face_tets.ensure(nt * 32 + 4));
    if (nt) {
        CK(cudaMemcpyAsync(ctx->tet_face_nodes.p, tet_face_nodes, nt * 48, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->tet_face_tets.p, tet_face_tets, nt * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->has_tet_faces = true;
    return sync_and_check(ctx);
}

int dscgpu_relabel_energies(dscgpu_ctx *ctx, int nb_phase, const double *mean, double alpha, double *energy)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && mean && energy, "bad arguments");
    if (!ctx->has_tet_faces) { ctx->err = "dscgpu_set_tet_faces has not been called for this snapshot"; return DSCGPU_ERR_NO_SNAPSHOT; }
    const size_t nt = ctx->n_tets;
    if (nt == 0) return DSCGPU_OK;
    CK(ctx->o_variation.ensure(nt * nb_phase * 8));
    // per-candidate sum |I - c_k| over the tet's sample points (get_variation), whole mesh regardless of the partition
    const uint32_t tb = ctx->tet_b, te = ctx->tet_e;
    ctx->tet_b = 0; ctx->tet_e = ctx->n_tets;
    int r = run_tet_pass(ctx, nb_phase, 0, nb_phase, mean, nullptr, nullptr, nullptr, nullptr, ctx->o_variation.as<double>(), ctx->sums.as<double>(),
                         ctx->mean.as<double>());
    ctx->tet_b = tb; ctx->tet_e = te;
    if (r) return r;
    const size_t n = nt * nb_phase;
    relabel_area_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->tet_face_nodes.as<unsigned>(), ctx->tet_face_tets.as<unsigned>(),
                                                                            nb_phase, alpha, ctx->o_variation.as<double>());
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(energy, ctx->o_variation.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_and_check(ctx);
}

int dscgpu_tet_sample_hash(dscgpu_ctx *ctx, uint64_t *tet_hash, int32_t *tet_nsamp)
{
    NEED_SNAPSHOT();
    ARG(tet_hash && tet_nsamp, "null output");
    const size_t nt = ctx->n_tets;
    if (nt == 0) return DSCGPU_OK;
    CK(ctx->o_hash.ensure(nt * 8));
    CK(ctx->o_nsamp.ensure(nt * 4));
    tet_sample_hash_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->d_tet_bary, ctx->o_hash.as<unsigned long long>(),
                                                                              ctx->o_nsamp.as<int>());
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tet_hash, ctx->o_hash.p, nt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(tet_nsamp, ctx->o_nsamp.p, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_and_check(ctx);
}

int dscgpu_points_in_tets(dscgpu_ctx *ctx, uint32_t n, const uint32_t *tet_idx, const double *pts, uint8_t *inside)
{
    NEED_SNAPSHOT();
    if (n == 0) return DSCGPU_OK;
    ARG(tet_idx && pts && inside, "null argument");
    for (uint32_t i = 0; i < n; i++) ARG(tet_idx[i] < ctx->n_tets, "tet index out of range");
    size_t bytes = (size_t)n * (4 + 24 + 1) + 64;
    CK(ctx->pit.ensure(bytes));
    double *d_pts = ctx->pit.as<double>();
    unsigned *d_idx = reinterpret_cast<unsigned *>(d_pts + 3 * (size_t)n);
    unsigned char *d_in = reinterpret_cast<unsigned char *>(d_idx + n);
    CK(cudaMemcpyAsync(d_pts, pts, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_idx, tet_idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    points_in_tets_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(mesh_view(ctx), n, d_idx, d_pts, d_in);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(inside, d_in, n, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_and_check(ctx);
}

int dscgpu_zray_means(dscgpu_ctx *ctx, int nb_phase, dscgpu_phase_stats *stats)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE, "nb_phase out of range");
    if (!ctx->has_sum_table) { ctx->err = "dscgpu_build_sum_table has not been called"; return DSCGPU_ERR_NO_SUMTABLE; }
    double h_sums[3 * kMaxPhase], h_mean[kMaxPhase];
    int r = run_zray_sync(ctx, nb_phase, 0, nullptr, ctx->sums.as<double>(), ctx->mean.as<double>(), h_sums, h_mean);
    if (r) return r;
    if (stats) fill_stats(stats, h_sums, h_mean, nb_phase);
    return DSCGPU_OK;
}

int dscgpu_vertex_bound(dscgpu_ctx *ctx, uint8_t *vertex_bound)
{
    NEED_SNAPSHOT();
    CK(ctx->vbound.ensure((size_t)ctx->n_nodes + 1));
    CK(cudaMemsetAsync(ctx->vbound.p, 0, (size_t)ctx->n_nodes + 1, ctx->stream));
    if (ctx->n_faces) {
        vertex_bound_kernel<<<(ctx->n_faces + 127) / 128, 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->vbound.as<unsigned char>());
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (vertex_bound) {
        CK(cudaMemcpyAsync(vertex_bound, ctx->vbound.p, ctx->n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DSCGPU_OK;
}

int dscgpu_zray_energy(dscgpu_ctx *ctx, int nb_phase, const double *mean, double alpha, const uint8_t *vertex_bound, double *energy_data,
                       double *energy_area)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && mean, "bad arguments");
    if (ctx->has_sum_table && !ctx->has_sq_table) { // the data term from two prefix tables instead of a loop over every interval
        const size_t n = (size_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2], XY = (size_t)ctx->dims[0] * ctx->dims[1];
        if (!ctx->d_sq_table) CK(cudaMalloc(&ctx->d_sq_table, n * sizeof(double)));
        const unsigned grid = (unsigned)((XY + 127) / 128);
        switch (ctx->dtype) {
        case DSCGPU_U8: sum_table_kernel<unsigned char, true><<<grid, 128, 0, ctx->stream>>>((const unsigned char *)ctx->d_vol, ctx->d_sq_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
        case DSCGPU_U16: sum_table_kernel<unsigned short, true><<<grid, 128, 0, ctx->stream>>>((const unsigned short *)ctx->d_vol, ctx->d_sq_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
        case DSCGPU_F32: sum_table_kernel<float, true><<<grid, 128, 0, ctx->stream>>>((const float *)ctx->d_vol, ctx->d_sq_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
        default: sum_table_kernel<double, true><<<grid, 128, 0, ctx->stream>>>((const double *)ctx->d_vol, ctx->d_sq_table, ctx->dims[0], ctx->dims[1], ctx->dims[2]); break;
        }
        ctx->launches++;
        CK(cudaGetLastError());
        ctx->has_sq_table = true;
    }
    double h_sums[3 * kMaxPhase];
    int r = run_zray_sync(ctx, nb_phase, 1, mean, ctx->sums.as<double>(), nullptr, h_sums, nullptr);
    if (r) return r;
    double d = 0;
    for (int k = 0; k < nb_phase; k++) d += h_sums[k]; // phases are accumulated one after another (:2449-2479)
    if (energy_data) *energy_data = d;
    // area term
    if (vertex_bound) {
        CK(ctx->vbound.ensure((size_t)ctx->n_nodes + 1));
        CK(cudaMemcpyAsync(ctx->vbound.p, vertex_bound, ctx->n_nodes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        r = dscgpu_vertex_bound(ctx, nullptr);
        if (r) return r;
    }
    double area = 0;
    if (ctx->n_faces) {
        unsigned grid = (ctx->n_faces + 255) / 256;
        CK(ctx->ray_psum.ensure(((size_t)grid + 1) * sizeof(double)));
        area_term_kernel<<<grid, 256, 0, ctx->stream>>>(mesh_view(ctx), ctx->vbound.as<unsigned char>(), alpha, ctx->ray_psum.as<double>());
        sum_partials_kernel<<<1, 32, 0, ctx->stream>>>(ctx->ray_psum.as<double>(), (int)grid, ctx->ray_psum.as<double>() + grid);
        ctx->launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&area, ctx->ray_psum.as<double>() + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (energy_area) *energy_area = area;
    return DSCGPU_OK;
}

int dscgpu_init_histogram(dscgpu_ctx *ctx, int nb_phase, int32_t hist[256])
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && hist, "bad arguments");
    const size_t nt = ctx->n_tets;
    memset(hist, 0, 256 * sizeof(int32_t));
    ctx->init_means_valid = false;
    if (!nt) { ctx->init_means_valid = true; return DSCGPU_OK; }
    CK(ctx->o_mean.ensure(nt * 8));
    CK(ctx->misc.ensure(256 * sizeof(int) + 64));
    // per-tet means of the whole mesh, whatever partition is set (the histogram is a property of the mesh), added in the reference's order:
    // the bins and the labels are truncations of these numbers
    {
        const unsigned g = (unsigned)((nt + 127) / 128);
        switch (ctx->dtype) {
        case DSCGPU_U8: tet_mean_ordered_kernel<unsigned char><<<g, 128, 0, ctx->stream>>>(vol_view(ctx), mesh_view(ctx), ctx->d_tet_bary, ctx->o_mean.as<double>()); break;
        case DSCGPU_U16: tet_mean_ordered_kernel<unsigned short><<<g, 128, 0, ctx->stream>>>(vol_view(ctx), mesh_view(ctx), ctx->d_tet_bary, ctx->o_mean.as<double>()); break;
        case DSCGPU_F32: tet_mean_ordered_kernel<float><<<g, 128, 0, ctx->stream>>>(vol_view(ctx), mesh_view(ctx), ctx->d_tet_bary, ctx->o_mean.as<double>()); break;
        default: tet_mean_ordered_kernel<double><<<g, 128, 0, ctx->stream>>>(vol_view(ctx), mesh_view(ctx), ctx->d_tet_bary, ctx->o_mean.as<double>()); break;
        }
        ctx->launches++;
        CK(cudaGetLastError());
    }
    int r;
    CK(cudaMemsetAsync(ctx->misc.p, 0, 256 * sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<size_t>((nt + 255) / 256, (size_t)ctx->num_sms * 8);
    init_hist_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->tet_label.as<int>(), ctx->o_mean.as<double>(), (unsigned)nt, ctx->misc.as<int>());
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hist, ctx->misc.p, 256 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    r = sync_and_check(ctx);
    if (r) return r;
    ctx->init_means_valid = true;
    return DSCGPU_OK;
}

int dscgpu_init_labels(dscgpu_ctx *ctx, int n_thresholds, const int32_t *thresholds, int32_t *tet_label_out)
{
    NEED_SNAPSHOT();
    ARG(n_thresholds >= 0 && n_thresholds < DSCGPU_MAX_PHASE && (n_thresholds == 0 || thresholds) && tet_label_out, "bad arguments");
    ARG(ctx->init_means_valid, "dscgpu_init_histogram must run on this snapshot first");
    const size_t nt = ctx->n_tets;
    if (!nt) return DSCGPU_OK;
    CK(ctx->misc.ensure(256 * sizeof(int) + 64));
    CK(ctx->o_nsamp.ensure(nt * sizeof(int)));
    if (n_thresholds) CK(cudaMemcpyAsync(ctx->misc.p, thresholds, n_thresholds * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    init_label_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->tet_label.as<int>(), ctx->o_mean.as<double>(), (unsigned)nt, n_thresholds,
                                                                          ctx->misc.as<int>(), ctx->o_nsamp.as<int>());
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tet_label_out, ctx->o_nsamp.p, nt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_and_check(ctx);
}

int dscgpu_set_node_face_order(dscgpu_ctx *ctx, uint32_t n_entries, const uint32_t *node_off, const uint32_t *face_idx)
{
    NEED_SNAPSHOT();
    if (!node_off) { ctx->has_face_order = false; return DSCGPU_OK; }
    ARG(n_entries == 0 || face_idx, "face_idx is required");
    const size_t nn = ctx->n_nodes;
    // checked on the host: a bad list would read outside the per-face arrays
    ARG(node_off[0] == 0 && node_off[nn] == n_entries, "node_off must start at 0 and end at n_entries");
    for (size_t i = 0; i < nn; i++) ARG(node_off[i] <= node_off[i + 1], "node_off must not decrease");
    for (uint32_t k = 0; k < n_entries; k++) ARG(face_idx[k] < ctx->n_faces, "face_idx out of range");
    CK(ctx->ord_off.ensure((nn + 1) * sizeof(unsigned)));
    CK(ctx->ord_idx.ensure(((size_t)n_entries + 1) * sizeof(unsigned)));
    CK(cudaMemcpyAsync(ctx->ord_off.p, node_off, (nn + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    if (n_entries) CK(cudaMemcpyAsync(ctx->ord_idx.p, face_idx, (size_t)n_entries * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    const int r = sync_and_check(ctx); // the host arrays may go away after the call
    ctx->has_face_order = r == DSCGPU_OK;
    return r;
}

int dscgpu_internal_forces(dscgpu_ctx *ctx, const uint8_t *node_flags, const uint8_t *vertex_bound, double *internal_forces)
{
    NEED_SNAPSHOT();
    ARG(node_flags && internal_forces, "node_flags and internal_forces are required");
    const size_t nn = ctx->n_nodes, nf = ctx->n_faces;
    int r;
    if (vertex_bound) {
        CK(ctx->vbound.ensure(nn + 1));
        CK(cudaMemcpyAsync(ctx->vbound.p, vertex_bound, nn, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        r = dscgpu_vertex_bound(ctx, nullptr);
        if (r) return r;
    }
    CK(ctx->nflags.ensure(nn + 1));
    CK(ctx->iforce.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->iterm.ensure(9 * nf * sizeof(double) + 8));
    CK(ctx->inorm.ensure(3 * nf * sizeof(double) + 8));
    CK(ctx->iskip.ensure(nf + 1));
    if (nn) CK(cudaMemcpyAsync(ctx->nflags.p, node_flags, nn, cudaMemcpyHostToDevice, ctx->stream));
    // nodes without interface faces keep a zero force (the reference's vector is zero-initialised, :2199)
    if (nn) CK(cudaMemsetAsync(ctx->iforce.p, 0, 3 * nn * sizeof(double), ctx->stream));
    if (nf) {
        iforce_face_kernel<<<(unsigned)((nf + 127) / 128), 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->vbound.as<unsigned char>(), ctx->iterm.as<double>(),
                                                                                ctx->inorm.as<double>(), ctx->iskip.as<unsigned char>());
        const unsigned max_active = std::min<unsigned>(ctx->n_nodes, 3u * ctx->n_faces);
        iforce_node_kernel<<<(max_active + 127) / 128, 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->node_active.as<unsigned>(), reinterpret_cast<const unsigned *>(ctr(ctx, 5)),
                                                                             ctx->nflags.as<unsigned char>(), ctx->iterm.as<double>(), ctx->inorm.as<double>(),
                                                                             ctx->iskip.as<unsigned char>(), ctx->has_face_order ? ctx->ord_off.as<unsigned>() : nullptr,
                                                                             ctx->has_face_order ? ctx->ord_idx.as<unsigned>() : nullptr, ctx->iforce.as<double>());
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    ctx->iforces_resident = true; // with the node flags and the vertex-bound flags this call used
    if (nn) CK(cudaMemcpyAsync(internal_forces, ctx->iforce.p, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_and_check(ctx);
}


int dscgpu_node_destinations(dscgpu_ctx *ctx, const dscgpu_move_params *mp, const double *forces, const double *internal_forces, const uint8_t *node_flags,
                             const uint8_t *vertex_bound, const double *dt_adapt, double *destinations, double *cur_dis, dscgpu_move_stats *stats)
{
    NEED_SNAPSHOT();
    ARG(mp && destinations, "params and destinations are required");
    const size_t nn = ctx->n_nodes;
    ARG(forces || ctx->forces_resident, "no external forces: pass them, or run dscgpu_image_force_eval / dscgpu_face_forces on this snapshot first");
    ARG(internal_forces || ctx->iforces_resident, "no internal forces: pass them, or run dscgpu_internal_forces on this snapshot first");
    ARG(node_flags || ctx->iforces_resident, "node_flags are required (or resident from dscgpu_internal_forces)");
    CK(ctx->forces.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->iforce.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->nflags.ensure(nn + 1));
    CK(ctx->vbound.ensure(nn + 1));
    CK(ctx->ndisp.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->ndest.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->misc.ensure(256 * sizeof(int) + 64));
    if (forces && nn) CK(cudaMemcpyAsync(ctx->forces.p, forces, 3 * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (internal_forces && nn) CK(cudaMemcpyAsync(ctx->iforce.p, internal_forces, 3 * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (node_flags && nn) CK(cudaMemcpyAsync(ctx->nflags.p, node_flags, nn, cudaMemcpyHostToDevice, ctx->stream));
    if (vertex_bound) {
        if (nn) CK(cudaMemcpyAsync(ctx->vbound.p, vertex_bound, nn, cudaMemcpyHostToDevice, ctx->stream));
    } else if (!ctx->iforces_resident) { // (dscgpu_internal_forces leaves the flags it used)
        int r = dscgpu_vertex_bound(ctx, nullptr);
        if (r) return r;
    }
    double *d_adapt = nullptr;
    if (dt_adapt && nn) {
        CK(ctx->iterm.ensure(nn * sizeof(double)));
        CK(cudaMemcpyAsync(ctx->iterm.p, dt_adapt, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        d_adapt = ctx->iterm.as<double>();
    }
    unsigned long long *maxbits = reinterpret_cast<unsigned long long *>(ctx->misc.p);
    CK(cudaMemsetAsync(maxbits, 0, 3 * sizeof(unsigned long long), ctx->stream));
    NodeMoveArgs a;
    a.pos4 = ctx->pos4.as<double4>();
    a.forces = ctx->forces.as<double>(); a.iforces = ctx->iforce.as<double>();
    a.vbound = ctx->vbound.as<unsigned char>(); a.node_flags = ctx->nflags.as<unsigned char>();
    a.dt_adapt = d_adapt;
    a.active_nodes = ctx->node_active.as<unsigned>(); a.n_active = reinterpret_cast<const unsigned *>(ctr(ctx, 5));
    a.alpha = mp->alpha; a.dt = mp->dt; a.threshold = mp->align_threshold; a.max_displacement = mp->max_displacement;
    for (int k = 0; k < 3; k++) a.dom[k] = (double)ctx->dims[k]; // image3d::dimension_v()
    a.disp = ctx->ndisp.as<double>(); a.dest = ctx->ndest.as<double>(); a.maxbits = maxbits; a.n_nodes = (unsigned)nn;
    if (nn) {
        node_dest_init_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(a.pos4, a.dest, a.disp, (unsigned)nn);
        const unsigned max_active = std::min<unsigned>(ctx->n_nodes, 3u * ctx->n_faces);
        const unsigned grid = std::max(1u, std::min<unsigned>((max_active + 255) / 256, (unsigned)ctx->num_sms * 4));
        node_displacement_kernel<<<grid, 256, 0, ctx->stream>>>(a);
        node_destination_kernel<<<grid, 256, 0, ctx->stream>>>(a);
        ctx->launches += 3;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(destinations, ctx->ndest.p, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (cur_dis) CK(cudaMemcpyAsync(cur_dis, ctx->ndisp.p, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    double h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, maxbits, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    int r = sync_and_check(ctx);
    if (r) return r;
    if (stats) {
        stats->max_dis = h[0]; stats->max_boundary_dis = h[1]; stats->real_max = h[2];
        stats->scale = h[0] > mp->max_displacement ? mp->max_displacement / h[0] : 1.0;
    }
    return DSCGPU_OK;
}

int dscgpu_thicken_forces(dscgpu_ctx *ctx, int nb_phase, const double *mean, const uint8_t *vertex_bound, double *forces, double *face_mean, double *face_dev,
                          uint8_t *face_evaluated)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && mean && forces, "bad arguments");
    ARG(ctx->face_b == 0 && ctx->face_e == ctx->n_faces && ctx->node_b == 0 && ctx->node_e == ctx->n_nodes, "dscgpu_thicken_forces evaluates the whole mesh: reset the partition");
    const size_t nn = ctx->n_nodes, nf = ctx->n_faces;
    CK(ctx->forces.ensure(3 * nn * sizeof(double) + 8));
    CK(ctx->vbound.ensure(nn + 1));
    CK(ctx->iskip.ensure(nf + 8));                     // masked Gauss-set indices
    CK(ctx->iterm.ensure(2 * nf * sizeof(double) + 8)); // per-face mean and deviation
    CK(cudaMemcpyAsync(ctx->mean.p, mean, nb_phase * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (vertex_bound) {
        if (nn) CK(cudaMemcpyAsync(ctx->vbound.p, vertex_bound, nn, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        int r = dscgpu_vertex_bound(ctx, nullptr);
        if (r) return r;
    }
    TimerScope ts(ctx, DSCGPU_T_FORCE);
    if (ctx->geom_async) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[19], 0));
        ctx->geom_async = false;
    } else {
        int rg = launch_force_geom(ctx, ctx->stream);
        if (rg) return rg;
    }
    CK(cudaMemsetAsync(ctx->forces.p, 0, 3 * nn * sizeof(double), ctx->stream));
    if (nf) {
        double *d_mu = ctx->iterm.as<double>(), *d_dev = d_mu + nf;
        face_force_stats_kernel<<<(unsigned)((nf + 127) / 128), 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->fgeo.as<double>(), ctx->flab.as<int>(), ctx->fint.as<double>(),
                                                                                    ctx->fset.as<signed char>(), ctx->vbound.as<unsigned char>(), nb_phase, ctx->mean.as<double>(),
                                                                                    ctx->iskip.as<signed char>(), d_mu, d_dev);
        ForceNode2Args na;
        na.mesh = mesh_view(ctx); na.fgeo = ctx->fgeo.as<double>(); na.flab = ctx->flab.as<int>(); na.fint = ctx->fint.as<double>();
        na.fset = ctx->iskip.as<signed char>(); na.nb_phase = nb_phase; na.mean = ctx->mean.as<double>(); na.forces = ctx->forces.as<double>();
        na.begin = 0; na.end = ctx->n_nodes;
        na.active_nodes = ctx->node_active.as<unsigned>();
        na.n_active = reinterpret_cast<const unsigned *>(ctr(ctx, 5));
        const unsigned max_active = std::min<unsigned>(ctx->n_nodes, 3u * ctx->n_faces);
        const unsigned grid = std::max(1u, std::min<unsigned>((max_active + 15) / 16, (unsigned)ctx->num_sms * 16));
        force_node2_kernel<<<grid, 128, 0, ctx->stream>>>(na);
        ctx->launches += 2;
        CK(cudaGetLastError());
        if (face_mean) CK(cudaMemcpyAsync(face_mean, d_mu, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (face_dev) CK(cudaMemcpyAsync(face_dev, d_dev, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (face_evaluated) CK(cudaMemcpyAsync(face_evaluated, ctx->iskip.p, nf, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (nn) CK(cudaMemcpyAsync(forces, ctx->forces.p, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int r = sync_and_check(ctx);
    if (r) return r;
    if (face_evaluated) for (size_t f = 0; f < nf; f++) face_evaluated[f] = ((signed char)face_evaluated[f]) >= 0 ? 1 : 0;
    ctx->forces_resident = true;
    return DSCGPU_OK;
}

int dscgpu_face_forces(dscgpu_ctx *ctx, int nb_phase, const double *mean, int mode, double *forces)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && mean && forces, "bad arguments");
    ARG(mode == DSCGPU_FORCE_GATHER || mode == DSCGPU_FORCE_ATOMIC, "unknown force mode");
    size_t bytes = sizeof(double) * 3 * (size_t)ctx->n_nodes;
    CK(ctx->forces.ensure(bytes + 8));
    CK(cudaMemcpyAsync(ctx->mean.p, mean, nb_phase * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int r = run_forces(ctx, nb_phase, ctx->mean.as<double>(), mode, ctx->forces.as<double>());
    ctx->forces_resident = r == DSCGPU_OK;
    if (r) return r;
    {
        TimerScope ts(ctx, DSCGPU_T_D2H);
        if (bytes) CK(cudaMemcpyAsync(forces, ctx->forces.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return sync_and_check(ctx);
}

int dscgpu_image_force_eval(dscgpu_ctx *ctx, const dscgpu_snapshot *snap, int nb_phase, int mean_mode, int force_mode, dscgpu_phase_stats *stats,
                            double *forces)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE, "nb_phase out of range");
    int r;
    bool tet_done = false;
    if (snap) {
        // stage the caller's arrays (host copy), upload below
        dscgpu_snapshot_staging st;
        ARG(snap->n_nodes_buffer == 0 || snap->pos, "pos is null");
        ARG(snap->n_tets == 0 || (snap->tet_nodes && snap->tet_label), "tet arrays are null");
        ARG(snap->n_faces == 0 || (snap->face_nodes && snap->face_tets && snap->face_is_boundary), "face arrays are null");
        r = dscgpu_snapshot_staging_get(ctx, snap->n_nodes_buffer, snap->n_tets, snap->n_faces, &st);
        if (r) return r;
        for (size_t i = 0; i < snap->n_nodes_buffer; i++) {
            st.pos4[4 * i] = snap->pos[3 * i]; st.pos4[4 * i + 1] = snap->pos[3 * i + 1]; st.pos4[4 * i + 2] = snap->pos[3 * i + 2]; st.pos4[4 * i + 3] = 0.0;
        }
        if (snap->n_tets) { memcpy(st.tet_nodes, snap->tet_nodes, (size_t)snap->n_tets * 16); memcpy(st.tet_label, snap->tet_label, (size_t)snap->n_tets * 4); }
        if (snap->n_faces) {
            memcpy(st.face_nodes, snap->face_nodes, (size_t)snap->n_faces * 12);
            memcpy(st.face_tets, snap->face_tets, (size_t)snap->n_faces * 8);
            memcpy(st.face_is_boundary, snap->face_is_boundary, (size_t)snap->n_faces);
        }
    }
    if (snap || ctx->staging_valid) {
        if (mean_mode == DSCGPU_MEAN_TET && ctx->pipeline_chunks > 0) {
            r = upload_and_tet_pass_pipelined(ctx, nb_phase, ctx->hot_want_sq, ctx->sums.as<double>(), ctx->mean.as<double>());
            if (r) return r;
            tet_done = true;
        } else {
            r = dscgpu_commit_snapshot(ctx);
            if (r) return r;
        }
    }
    NEED_SNAPSHOT();
    size_t fbytes = sizeof(double) * 3 * (size_t)ctx->n_nodes;
    CK(ctx->forces.ensure(fbytes + 8));
    CK(ctx->h_out.ensure(fbytes + 4 * kMaxPhase * sizeof(double)));
    if (mean_mode == DSCGPU_MEAN_ZRAY) {
        if (!ctx->has_sum_table) { ctx->err = "dscgpu_build_sum_table has not been called"; return DSCGPU_ERR_NO_SUMTABLE; }
        r = run_zray_sync(ctx, nb_phase, 0, nullptr, ctx->sums.as<double>(), ctx->mean.as<double>(), nullptr, nullptr);
    } else if (!tet_done) {
        r = run_tet_pass(ctx, nb_phase, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->sums.as<double>(), ctx->mean.as<double>(), ctx->hot_want_sq, true);
    } else {
        r = DSCGPU_OK;
    }
    if (r) return r;
    r = run_forces(ctx, nb_phase, ctx->mean.as<double>(), force_mode, ctx->forces.as<double>());
    ctx->forces_resident = r == DSCGPU_OK;
    if (r) return r;
    trace_mark(ctx, "forces done", ctx->stream);
    double *h = ctx->h_out.as<double>();
    {
        TimerScope ts(ctx, DSCGPU_T_D2H);
        const bool compact = fbytes && !forces && ctx->compact_forces && ctx->ev_nact;
        if (!compact) {
            CK(cudaMemcpyAsync(h, ctx->sums.p, 3 * nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(h + 3 * kMaxPhase, ctx->mean.p, nb_phase * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (compact) {
            // only the rows that can be non-zero cross PCIe: nodes with interface faces (about one node in seven on a DSC mesh)
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            ARG(cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone, "the compact read-back cannot be captured");
            CK(cudaEventSynchronize(ctx->ev_nact)); // recorded when the snapshot was committed
            const unsigned na = std::min<unsigned>(*ctx->h_nact.as<unsigned>(), ctx->n_nodes);
            // statistics, rows and keys in one buffer: one copy (every copy costs a launch on the tail of the evaluation)
            const size_t bytes = (4 * (size_t)kMaxPhase + 3 * (size_t)na) * sizeof(double) + (size_t)na * sizeof(unsigned);
            CK(ctx->crows.ensure(bytes + 8));
            CK(ctx->h_out.ensure(std::max(bytes + 8, fbytes + 4 * kMaxPhase * sizeof(double))));
            h = ctx->h_out.as<double>();
            gather_result_kernel<<<(std::max<unsigned>(na, 3 * kMaxPhase) + 255) / 256, 256, 0, ctx->stream>>>(ctx->forces.as<double>(), ctx->node_active.as<unsigned>(), na,
                                                                                                              ctx->sums.as<double>(), ctx->mean.as<double>(), nb_phase,
                                                                                                              ctx->crows.as<double>());
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(h, ctx->crows.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->compact_n = na;
            ctx->compact_keys = reinterpret_cast<const uint32_t *>(h + 4 * kMaxPhase + 3 * (size_t)na);
        } else if (fbytes) {
            CK(cudaMemcpyAsync(h + 4 * kMaxPhase, ctx->forces.p, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->compact_n = 0xffffffffu;
        }
    }
    trace_mark(ctx, "results on the host", ctx->stream);
    r = sync_and_check(ctx);
    trace_dump(ctx);
    if (r) return r;
    if (stats) fill_stats(stats, h, h + 3 * kMaxPhase, nb_phase);
    if (forces && fbytes) memcpy(forces, h + 4 * kMaxPhase, fbytes);
    return DSCGPU_OK;
}

int dscgpu_set_overlap_geometry(dscgpu_ctx *ctx, int on)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->overlap_geom = on ? 1 : 0;
    return DSCGPU_OK;
}

int dscgpu_set_compact_forces(dscgpu_ctx *ctx, int on)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->compact_forces = on ? 1 : 0;
    return DSCGPU_OK;
}

int dscgpu_read_forces_compact(dscgpu_ctx *ctx, const double *d_forces, uint32_t *n_rows, const uint32_t **node_keys, const double **rows)
{
    NEED_SNAPSHOT();
    ARG(d_forces && n_rows && node_keys && rows, "null argument");
    ARG(ctx->ev_nact, "no node -> face table on this snapshot");
    CK(cudaEventSynchronize(ctx->ev_nact)); // recorded when the snapshot was committed
    const unsigned na = std::min<unsigned>(*ctx->h_nact.as<unsigned>(), ctx->n_nodes);
    CK(ctx->crows.ensure(3 * (size_t)na * sizeof(double) + 8));
    CK(ctx->h_node_active.ensure((size_t)na * sizeof(unsigned) + 8));
    CK(ctx->h_out.ensure(3 * (size_t)na * sizeof(double) + 4 * kMaxPhase * sizeof(double)));
    double *h = ctx->h_out.as<double>() + 4 * kMaxPhase;
    if (na) {
        TimerScope ts(ctx, DSCGPU_T_D2H);
        gather_rows_kernel<<<(na + 255) / 256, 256, 0, ctx->stream>>>(d_forces, ctx->node_active.as<unsigned>(), na, ctx->crows.as<double>());
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h, ctx->crows.p, 3 * (size_t)na * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_node_active.p, ctx->node_active.p, (size_t)na * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    }
    int r = sync_and_check(ctx);
    if (r) return r;
    ctx->compact_n = na;
    ctx->compact_keys = ctx->h_node_active.as<uint32_t>();
    *n_rows = na; *node_keys = ctx->compact_keys; *rows = h;
    return DSCGPU_OK;
}

int dscgpu_forces_compact_pinned(dscgpu_ctx *ctx, uint32_t *n_rows, const uint32_t **node_keys, const double **rows)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(n_rows && node_keys && rows, "null output");
    ARG(ctx->h_out.p && ctx->compact_n != 0xffffffffu, "the last dscgpu_image_force_eval did not read the forces back in compact form");
    *n_rows = ctx->compact_n;
    *node_keys = ctx->compact_keys;
    *rows = ctx->h_out.as<double>() + 4 * kMaxPhase;
    return DSCGPU_OK;
}

const double *dscgpu_forces_pinned(dscgpu_ctx *ctx, uint32_t *n_nodes_buffer)
{
    if (!ctx || !ctx->h_out.p || ctx->compact_n != 0xffffffffu) return nullptr; // (compact read-back: dscgpu_forces_compact_pinned)
    if (n_nodes_buffer) *n_nodes_buffer = ctx->n_nodes;
    return ctx->h_out.as<double>() + 4 * kMaxPhase;
}

int dscgpu_tet_phase_sums_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_sums, "bad arguments");
    return run_tet_pass(ctx, nb_phase, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_sums, nullptr, ctx->hot_want_sq, true);
}

int dscgpu_zray_phase_sums_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums)
{
    NEED_SNAPSHOT();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_sums, "bad arguments");
    if (!ctx->has_sum_table) { ctx->err = "dscgpu_build_sum_table has not been called"; return DSCGPU_ERR_NO_SUMTABLE; }
    // under stream capture the pass is recorded as it is (no host round trip): run it once outside capture first, which sizes the hit buffer
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return run_zray(ctx, nb_phase, 0, nullptr, d_sums, nullptr);
    return run_zray_sync(ctx, nb_phase, 0, nullptr, d_sums, nullptr, nullptr, nullptr);
}

int dscgpu_means_from_sums_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_sums, double *d_mean)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_sums && d_mean, "bad arguments");
    means_from_sums_kernel<<<1, 32, 0, ctx->stream>>>(d_sums, nb_phase, d_mean);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

int dscgpu_face_geometry_async(dscgpu_ctx *ctx)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    // fork: the side stream picks up where the main stream is now (snapshot in place), runs the kernel, and the next
    // gather-mode force call joins it; both edges are events, so the pattern can be captured in a CUDA graph
    CK(cudaEventRecord(ctx->ev_chunk[18], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[18], 0));
    int r = launch_force_geom(ctx, ctx->copy_stream);
    if (r) return r;
    CK(cudaEventRecord(ctx->ev_chunk[19], ctx->copy_stream));
    ctx->geom_async = true;
    return DSCGPU_OK;
}

int dscgpu_face_forces_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_mean, int mode, double *d_forces)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_mean && d_forces, "bad arguments");
    return run_forces(ctx, nb_phase, d_mean, mode, d_forces);
}

// ---- peer-memory exchange (dsc_comm.cuh) -------------------------------------------------------------------------------

int dscgpu_comm_create(dscgpu_ctx *ctx, int rank, int world, uint32_t cap_rows, unsigned char handle_out[64])
{
    return dscgpu_comm_create_ex(ctx, rank, world, cap_rows, 0, handle_out);
}

int dscgpu_comm_create_ex(dscgpu_ctx *ctx, int rank, int world, uint32_t cap_rows, uint64_t snapshot_bytes, unsigned char handle_out[64])
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "rank / world out of range (at most 8 ranks)");
    ARG(cap_rows > 0, "cap_rows must be positive");
    ARG(!ctx->comm_win[rank] && ctx->comm_world == 0, "the exchange window exists already");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const size_t snap = ((size_t)snapshot_bytes + 255) & ~(size_t)255;
    const size_t bytes = comm_window_bytes(world, cap_rows, snap);
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemset(p, 0, comm_snapshot_offset(world, cap_rows)));
    ctx->comm_snap_bytes = snap;
    {   // scratch of the fused force exchange, both halves: barrier counters 0, touched range empty
        const unsigned init[8] = {0u, 0u, 0xffffffffu, 0u, 0u, 0u, 0xffffffffu, 0u};
        CK(cudaMemcpy(static_cast<unsigned char *>(p) + kCommLocal, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&ctx->comm_rows, 3 * (size_t)cap_rows * sizeof(double)));
    CK(cudaMemset(ctx->comm_rows, 0, 3 * (size_t)cap_rows * sizeof(double)));
    ctx->comm_win[rank] = static_cast<unsigned char *>(p);
    ctx->comm_rank = rank; ctx->comm_world = world; ctx->comm_cap_rows = cap_rows;
    ctx->comm_connected = world == 1;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, p));
        memcpy(handle_out, &h, 64);
    }
    return DSCGPU_OK;
}

int dscgpu_comm_connect(dscgpu_ctx *ctx, const unsigned char *handles)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(ctx->comm_world >= 1 && handles, "dscgpu_comm_create first");
    for (int k = 0; k < ctx->comm_world; k++) {
        if (k == ctx->comm_rank || ctx->comm_win[k]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)k, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->comm_win[k] = static_cast<unsigned char *>(p);
        ctx->comm_ipc[k] = true;
    }
    ctx->comm_connected = true;
    return DSCGPU_OK;
}

int dscgpu_comm_connect_local(dscgpu_ctx *ctx, void *const *windows)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(ctx->comm_world >= 1 && windows, "dscgpu_comm_create first");
    for (int k = 0; k < ctx->comm_world; k++)
        if (k != ctx->comm_rank) ctx->comm_win[k] = static_cast<unsigned char *>(windows[k]);
    ctx->comm_connected = true;
    return DSCGPU_OK;
}

void *dscgpu_comm_window_dev(dscgpu_ctx *ctx) { return (ctx && ctx->comm_world) ? ctx->comm_win[ctx->comm_rank] : nullptr; }

int dscgpu_comm_error(dscgpu_ctx *ctx)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(ctx->comm_world >= 1, "no exchange window");
    int e = 0;
    CK(cudaMemcpyAsync(&e, ctx->comm_win[ctx->comm_rank] + kCommError, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (e) { // reported once: the word is cleared, the results of that exchange are NaN
        CK(cudaMemsetAsync(ctx->comm_win[ctx->comm_rank] + kCommError, 0, sizeof(int), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->err = "peer exchange timed out: a rank did not reach the exchange within 2 s";
        return DSCGPU_ERR_CUDA;
    }
    return DSCGPU_OK;
}

int dscgpu_tet_phase_sums_allreduce_dev(dscgpu_ctx *ctx, int nb_phase, double *d_sums, double *d_mean)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_sums, "bad arguments");
    ARG(ctx->comm_connected, "dscgpu_comm_create / dscgpu_comm_connect first");
    int r = tet_pass_begin(ctx);
    if (r) return r;
    TetPassOut none;
    {
        TimerScope ts(ctx, DSCGPU_T_TET);
        r = tet_pass_range(ctx, nb_phase, 0, 0, nullptr, none, ctx->tet_b, ctx->tet_e, ctx->hot_want_sq);
        if (r) return r;
        r = launch_exact_list(ctx);
        if (r) return r;
    }
    // the mean-independent half of the force pass starts on the side stream now: it runs while this rank waits for the sums of the others
    r = dscgpu_face_geometry_async(ctx);
    if (r) return r;
    TimerScope ts(ctx, DSCGPU_T_PHASE_FINAL);
    if (ctx->tet_rows_fixed)
        comm_phase_allreduce_kernel<true><<<1, 32 * 3 * kMaxPhase, 0, ctx->stream>>>(comm_view(ctx), ctx->d_stg_counter + 4, 1, nb_phase, ctx->stg_scale[0], ctx->stg_scale[1],
                                                                                     ctx->stg_scale[2], d_sums, d_mean, ctx->d_stg_counter);
    else
        comm_phase_allreduce_kernel<false><<<1, 32 * 3 * kMaxPhase, 0, ctx->stream>>>(comm_view(ctx), ctx->partials.p, ctx->tet_grid, nb_phase, 1.0, 1.0, 1.0, d_sums, d_mean, nullptr);
    ctx->launches++;
    ctx->stg_pass_open = false;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

int dscgpu_face_forces_allreduce_dev(dscgpu_ctx *ctx, int nb_phase, const double *d_mean, double *d_forces)
{
    NEED_SNAPSHOT();
    NEED_PARTITION();
    ARG(nb_phase >= 1 && nb_phase <= DSCGPU_MAX_PHASE && d_mean && d_forces, "bad arguments");
    ARG(ctx->comm_connected, "dscgpu_comm_create / dscgpu_comm_connect first");
    ARG(ctx->n_nodes <= ctx->comm_cap_rows, "the exchange window is too small for this snapshot (cap_rows < n_nodes_buffer)");
    TimerScope ts(ctx, DSCGPU_T_FORCE);
    // mean-independent half (normals, areas, trilinear samples of this rank's faces): it may already be running on the side stream,
    // started before the sums exchange (dscgpu_face_geometry_async)
    if (ctx->geom_async) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[19], 0));
        ctx->geom_async = false;
    } else {
        int rg = launch_force_geom(ctx, ctx->stream);
        if (rg) return rg;
    }
    ForceFusedArgs a;
    a.c = comm_view(ctx);
    a.mesh = mesh_view(ctx);
    a.fgeo = ctx->fgeo.as<double>(); a.flab = ctx->flab.as<int>(); a.fint = ctx->fint.as<double>(); a.fset = ctx->fset.as<signed char>();
    a.nb_phase = nb_phase;
    a.mean = d_mean;
    a.apos = ctx->csr_apos.as<unsigned>();
    a.rows = ctx->comm_rows;
    a.begin = ctx->face_b; a.end = ctx->face_e;
    a.active_nodes = ctx->node_active.as<unsigned>();
    a.n_active_ptr = reinterpret_cast<const unsigned *>(ctr(ctx, 5));
    a.forces = d_forces;
    a.n_nodes = ctx->n_nodes;
    // one launch: scatter, push of the touched rows, wait, ordered sum.  The grid spins on its own barrier: it must be co-resident
    static int occ = 0;
    if (occ == 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, comm_force_fused_kernel, 256, 0));
        if (occ < 1) occ = 1;
        if (occ > 2) occ = 2;
    }
    unsigned grid = (unsigned)(ctx->num_sms * occ);
    const unsigned want = (std::max<unsigned>(std::max<unsigned>(a.end - a.begin, ctx->n_nodes / 8), 256u) + 255) / 256;
    if (want < grid) grid = want;
    comm_force_fused_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSCGPU_OK;
}

int dscgpu_enable_timers(dscgpu_ctx *ctx, int on)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->timers_on = on != 0;
    return DSCGPU_OK;
}

int dscgpu_get_timer(dscgpu_ctx *ctx, int which, float *ms)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(which >= 0 && which < DSCGPU_T_COUNT && ms, "bad timer");
    *ms = 0;
    if (!ctx->ev_used[which]) return DSCGPU_OK;
    CK(cudaEventSynchronize(ctx->ev[which][1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
    return DSCGPU_OK;
}

int dscgpu_set_tet_variant(dscgpu_ctx *ctx, int variant)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(variant >= 0 && (variant & ~(128 | 4096 | 16384 | 131072)) == 0, "unknown tet-kernel variant bit (dscgpu.h lists them)");
    ctx->tet_variant = variant;
    return DSCGPU_OK;
}

int dscgpu_get_tet_variant(const dscgpu_ctx *ctx) { return ctx ? ctx->tet_variant : -1; }

int dscgpu_set_hot_want_sq(dscgpu_ctx *ctx, int on)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ctx->hot_want_sq = on ? 1 : 0;
    return DSCGPU_OK;
}

uint64_t dscgpu_launch_count(const dscgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int dscgpu_set_lanes_per_tet(dscgpu_ctx *ctx, int lanes)
{
    if (!ctx) return DSCGPU_ERR_ARG;
    ARG(lanes == 0 || lanes == 8 || lanes == 32, "lanes per tet must be 0, 8 or 32");
    ctx->lanes_per_tet = lanes;
    return DSCGPU_OK;
}

static uint64_t read_counter(dscgpu_ctx *ctx, int i)
{
    unsigned long long v = 0;
    if (!ctx || !ctx->counters.p) return 0;
    cudaMemcpyAsync(&v, ctr(ctx, i), sizeof(v), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return v;
}
uint64_t dscgpu_last_sample_count(dscgpu_ctx *ctx) { return read_counter(ctx, 0); }
uint64_t dscgpu_last_gauss_count(dscgpu_ctx *ctx)
{
    if (!ctx || !ctx->has_snapshot || !ctx->counters.p) return 0;
    cudaMemsetAsync(ctr(ctx, 1), 0, sizeof(unsigned long long), ctx->stream);
    const unsigned n = ctx->face_e - ctx->face_b;
    if (n) {
        gauss_count_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(mesh_view(ctx), ctx->face_b, ctx->face_e, ctr(ctx, 1));
        ctx->launches++;
    }
    return read_counter(ctx, 1);
}

int dscgpu_host_sort_hits(int n, int32_t *z, uint8_t *b_in)
{
    if (n < 0 || (n > 0 && (!z || !b_in))) return DSCGPU_ERR_ARG;
    std::vector<int> v(n);
    for (int i = 0; i < n; i++) v[i] = (int)((unsigned)z[i] << 1) | (b_in[i] ? 1 : 0);
    if (n > 0) std_sort_hits(v.data(), n);
    for (int i = 0; i < n; i++) { z[i] = v[i] >> 1; b_in[i] = (uint8_t)(v[i] & 1); }
    return DSCGPU_OK;
}

} // extern "C"
