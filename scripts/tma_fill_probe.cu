// Probe: how fast can 8 warps per SM fill private shared-memory slots with C4-like voxel boxes, and with which instruction?
//   mode 0  one cp.async.bulk (1-D, one row of the box) per (y,z) row, issued by all 32 lanes
//   mode 1  one cp.async.bulk.tensor.3d per z slice (box ex x ey x 1), lanes 0..ez-1
//   mode 2  one cp.async.bulk.tensor.3d per tile (box ex x ey x ez), lane 0
//   mode 3  LDG.128 + STS.128 by the warp (no TMA), for comparison
// Every warp loops over tiles (dynamic fetch), waits for its box, optionally "computes" (work FFMA iterations plus LDS
// gathers) and goes on.  Prints ms, GB/s of shared-memory fill and tiles/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tma_fill_probe.bin scripts/tma_fill_probe.cu
//   scripts/tma_fill_probe.bin MODE [work] [warps] [ex ey ez] [prefetch_l2]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Params {
    const float *vol;
    int X, Y, Z;
    int ex, ey, ez;      // box (elements)
    int tx, ty, tz;      // tiles per axis
    int sx, sy, sz;      // tile pitch (elements)
    int n_tiles;
    int work;
    int mode;
    int slot_bytes;
    int prefetch;
    unsigned *counter;
    float *sink;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, const Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char *slot = smem + (size_t)warp * p.slot_bytes;
    const unsigned slot_a = smem_u32(slot), bar_a = smem_u32(&bars[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    float acc = 0.f;
    const int rows = p.ey * p.ez, row_bytes = p.ex * 4, box_bytes = rows * row_bytes;
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(p.counter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < (unsigned)p.n_tiles) {
        const int ix = tile % p.tx, iy = (tile / p.tx) % p.ty, iz = tile / (p.tx * p.ty);
        const int x0 = ix * p.sx, y0 = iy * p.sy, z0 = iz * p.sz;
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(p.counter, 1u);
        if (p.mode == 3) {
            const int vec_per_row = p.ex / 4;
            for (int i = lane; i < rows * vec_per_row; i += 32) {
                const int r = i / vec_per_row, c = i - r * vec_per_row;
                const int y = y0 + r % p.ey, z = z0 + r / p.ey;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p.vol + ((size_t)z * p.Y + y) * p.X + x0) + c);
                reinterpret_cast<float4 *>(slot)[i] = v;
            }
            __syncwarp();
        } else {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(box_bytes) : "memory");
            __syncwarp();
            if (p.mode == 0) {
                for (int r = lane; r < rows; r += 32) {
                    const int y = y0 + r % p.ey, z = z0 + r / p.ey;
                    const float *src = p.vol + ((size_t)z * p.Y + y) * p.X + x0;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot_a + r * row_bytes), "l"(src),
                                 "r"(row_bytes), "r"(bar_a)
                                 : "memory");
                }
            } else if (p.mode == 1) {
                if (lane < p.ez)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                                     slot_a + lane * p.ey * row_bytes),
                                 "l"(&tm), "r"(x0), "r"(y0), "r"(z0 + lane), "r"(bar_a)
                                 : "memory");
            } else {
                if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(slot_a), "l"(&tm),
                                 "r"(x0), "r"(y0), "r"(z0), "r"(bar_a)
                                 : "memory");
            }
            unsigned done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            phase ^= 1;
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        if (p.prefetch && next < (unsigned)p.n_tiles && lane < p.ez && p.mode != 3) {
            const int jx = next % p.tx, jy = (next / p.tx) % p.ty, jz = next / (p.tx * p.ty);
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&tm), "r"(jx * p.sx), "r"(jy * p.sy), "r"(jz * p.sz + (p.mode == 2 ? 0 : lane))
                         : "memory");
        }
        // stand-in for the sample loop: work x (a pseudo-random LDS gather + a few FMAs)
        const float *box = reinterpret_cast<const float *>(slot);
        const int n_el = box_bytes / 4;
        unsigned h = lane * 2654435761u + tile;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int i = 0; i < p.work; i++) {
            h = h * 1664525u + 1013904223u;
            const unsigned i0 = (h >> 8) % (unsigned)n_el, i1 = (h >> 12) % (unsigned)n_el, i2 = (h >> 16) % (unsigned)n_el, i3 = (h >> 4) % (unsigned)n_el;
            a0 = fmaf(box[i0], 1.0001f, a0); a1 = fmaf(box[i1], 1.0002f, a1); a2 = fmaf(box[i2], 1.0003f, a2); a3 = fmaf(box[i3], 1.0004f, a3);
        }
        acc += (a0 + a1) + (a2 + a3) + box[(lane * 37) % n_el];
        __syncwarp();
        tile = next;
    }
    if (acc == 123.456f) p.sink[threadIdx.x] = acc;
    (void)nw;
}

int main(int argc, char **argv)
{
    Params p{};
    p.mode = argc > 1 ? atoi(argv[1]) : 0;
    p.work = argc > 2 ? atoi(argv[2]) : 0;
    const int warps = argc > 3 ? atoi(argv[3]) : 8;
    p.ex = argc > 4 ? atoi(argv[4]) : 52; p.ey = argc > 5 ? atoi(argv[5]) : 10; p.ez = argc > 6 ? atoi(argv[6]) : 10;
    p.prefetch = argc > 7 ? atoi(argv[7]) : 0;
    p.X = p.Y = p.Z = 512;
    p.sx = 40; p.sy = 7; p.sz = 7;
    p.tx = (p.X - p.ex) / p.sx + 1; p.ty = (p.Y - p.ey) / p.sy + 1; p.tz = (p.Z - p.ez) / p.sz + 1;
    p.n_tiles = p.tx * p.ty * p.tz;
    p.slot_bytes = ((p.ex * p.ey * p.ez * 4 + 127) / 128) * 128;
    const size_t nvox = (size_t)p.X * p.Y * p.Z;
    float *vol;
    CK(cudaMalloc(&vol, nvox * 4));
    CK(cudaMemset(vol, 0, nvox * 4));
    p.vol = vol;
    CK(cudaMalloc(&p.counter, 4));
    CK(cudaMalloc(&p.sink, 4096));
    void *flush;
    CK(cudaMalloc(&flush, 512u << 20));

    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        const cuuint64_t gdim[3] = {(cuuint64_t)p.X, (cuuint64_t)p.Y, (cuuint64_t)p.Z};
        const cuuint64_t gstr[2] = {(cuuint64_t)p.X * 4, (cuuint64_t)p.X * p.Y * 4};
        const cuuint32_t box[3] = {(cuuint32_t)p.ex, (cuuint32_t)p.ey, (cuuint32_t)(p.mode == 2 ? p.ez : 1)};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled -> %d\n", (int)r); return 1; }
    }
    const int smem = warps * p.slot_bytes;
    CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev_sms = 0;
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 5; it++) {
        CK(cudaMemsetAsync(flush, it, 512u << 20));
        CK(cudaMemsetAsync(p.counter, 0, 4));
        cudaEventRecord(e0);
        fill_kernel<<<dev_sms, warps * 32, smem>>>(tm, p);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    const double fill = (double)p.n_tiles * p.ex * p.ey * p.ez * 4;
    printf("mode %d work %d warps %d box %dx%dx%d pf %d: tiles %d  %.4f ms  fill %.3f GB -> %.0f GB/s  (%.2f Mtiles/s)\n", p.mode, p.work, warps, p.ex, p.ey, p.ez, p.prefetch,
           p.n_tiles, best, fill / 1e9, fill / best / 1e6, p.n_tiles / best / 1e3);
    return 0;
}
