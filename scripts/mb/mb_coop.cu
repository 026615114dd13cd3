// Micro-benchmark (measurement tool, not product): what does a random access cost past the reach of the first-level
// TLB (256 MB) -- per lane, per sector, or per (instruction, 128-byte line)?  Variants, all random over a contiguous
// footprint, rates in accesses (sectors / lines as named) per second:
//   lane32      every lane its own random 32-byte sector, one 256-bit load            (what a sliced row load does today)
//   lane128x4   every lane its own random 128-byte line, four 256-bit loads          (four instructions per line)
//   coop2x32    two adjacent lanes share a random 64-byte half line, one 256-bit load each
//   coop4x32    four adjacent lanes share a random 128-byte line, one 256-bit load each (ONE instruction per line)
//   coop8x16    eight adjacent lanes share a line, one 128-bit load each
//   coop32x4    the whole warp reads one line, 4 bytes per lane
//   tma128/32   every lane issues cp.async.bulk global -> shared of 128 / 32 bytes, completion on an mbarrier
// Output: CSV  variant,footprint_MB,G_access_per_s,GB_per_s
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld256(const void *p) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld128(const void *p) {
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}
__device__ __forceinline__ uint32_t ld32(const void *p) {
    uint32_t a;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a) : "l"(p));
    return a;
}

enum { LANE32, LANE128X4, COOP2, COOP4, COOP8, COOP32 };

// COOP = lanes sharing one random target; the generator is seeded by the group so that its lanes agree
template <int MODE, int UNROLL>
__global__ void gather(const uint8_t *__restrict__ buf, uint64_t n_lines, int iters, uint32_t *sink) {
    constexpr int COOP = MODE == COOP2 ? 2 : MODE == COOP4 ? 4 : MODE == COOP8 ? 8 : MODE == COOP32 ? 32 : 1;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, sub = threadIdx.x % COOP;
    uint64_t x = (uint64_t)(tid / COOP) * 0x9e3779b97f4a7c15ULL + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            const uint64_t u = (uint64_t)(((unsigned __int128)x * n_lines) >> 64);
            const uint8_t *p = buf + u * 128;
            if (MODE == LANE32) acc ^= ld256(p + (x & 3) * 32);
            else if (MODE == LANE128X4) {
#pragma unroll
                for (int w = 0; w < 4; ++w) acc ^= ld256(p + w * 32);
            } else if (MODE == COOP2) acc ^= ld256(p + (x & 1) * 64 + sub * 32);
            else if (MODE == COOP4) acc ^= ld256(p + sub * 32);
            else if (MODE == COOP8) acc ^= ld128(p + sub * 16);
            else acc ^= ld32(p + sub * 4);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ---- bulk-copy (TMA) variant: 4 warps per CTA, every lane keeps UN requests of BYTES in flight
template <int BYTES, int UN>
__global__ void __launch_bounds__(128) gather_tma(const uint8_t *__restrict__ buf, uint64_t n_lines, int iters, uint32_t *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
    uint8_t *mine = smem + (size_t)warp * 32 * UN * BYTES;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(mine) + lane * UN * BYTES;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL + 12345;
    uint32_t acc = 0, parity = 0;
    for (int i = 0; i < iters; ++i) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * UN * BYTES) : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < UN; ++j) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            const uint64_t u = (uint64_t)(((unsigned __int128)x * n_lines) >> 64);
            const uint8_t *p = buf + u * 128 + (BYTES == 32 ? (x & 3) * 32 : 0);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + j * BYTES),
                         "l"(p), "r"(BYTES), "r"(bar)
                         : "memory");
        }
        uint32_t ok = 0;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
        } while (!ok);
        parity ^= 1;
        acc ^= *(const volatile uint32_t *)(mine + lane * UN * BYTES);
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static double timed(F launch, double accesses_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 40;
    launch(20);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < 10.f && rep == 0 && iters < 2560) { iters *= 4; rep = -1; continue; }
        const double rate = accesses_per_iter * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

template <int MODE>
static void run(const char *name, const uint8_t *buf, uint64_t bytes, uint32_t *sink, int bytes_per_access) {
    constexpr int COOP = MODE == COOP2 ? 2 : MODE == COOP4 ? 4 : MODE == COOP8 ? 8 : MODE == COOP32 ? 32 : 1;
    constexpr int UN = 10;
    const int grid = 148 * 4, block = 256;
    const uint64_t n_lines = bytes / 128;
    const double r = timed([&](int it) { gather<MODE, UN><<<grid, block>>>(buf, n_lines, it, sink); }, (double)grid * block / COOP * UN);
    printf("%s,%.0f,%.2f,%.1f\n", name, bytes / 1048576.0, r / 1e9, r * bytes_per_access / 1e9);
    fflush(stdout);
}

// the same quad-of-lanes line gather with fewer loads in flight per lane and fewer CTAs per SM (what a kernel with real
// work between the loads can afford)
template <int UN>
static void run_depth(const uint8_t *buf, uint64_t bytes, uint32_t *sink, int ctas_per_sm) {
    const int grid = 148 * ctas_per_sm, block = 256;
    const uint64_t n_lines = bytes / 128;
    const double r = timed([&](int it) { gather<COOP4, UN><<<grid, block>>>(buf, n_lines, it, sink); }, (double)grid * block / 4 * UN);
    printf("coop4x32_inflight%d_ctas%d,%.0f,%.2f,%.1f\n", UN, ctas_per_sm, bytes / 1048576.0, r / 1e9, r * 128 / 1e9);
    fflush(stdout);
}

template <int BYTES>
static void run_tma(const char *name, const uint8_t *buf, uint64_t bytes, uint32_t *sink) {
    constexpr int UN = BYTES == 128 ? 4 : 8;
    const int grid = 148 * 3, block = 128;
    const size_t smem = (size_t)4 * 32 * UN * BYTES;
    cudaFuncSetAttribute(gather_tma<BYTES, UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t n_lines = bytes / 128;
    const double r = timed([&](int it) { gather_tma<BYTES, UN><<<grid, block, smem>>>(buf, n_lines, it, sink); }, (double)grid * block * UN);
    printf("%s,%.0f,%.2f,%.1f\n", name, bytes / 1048576.0, r / 1e9, r * BYTES / 1e9);
    fflush(stdout);
}

int main(int argc, char **) {
    const bool depth_only = argc > 1;
    uint32_t *sink = nullptr;
    cudaMalloc(&sink, 4);
    const uint64_t big = 16ull << 30;
    uint8_t *buf = nullptr;
    if (cudaMalloc(&buf, big) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemset(buf, 0x5a, big);
    printf("variant,footprint_MB,G_access_per_s,GB_per_s\n");
    for (uint64_t mb : {230ull, 460ull, 1840ull, 16384ull}) {
        if (depth_only) break;
        const uint64_t bytes = mb << 20;
        run<LANE32>("lane32", buf, bytes, sink, 32);
        run<LANE128X4>("lane128x4", buf, bytes, sink, 128);
        run<COOP2>("coop2x32", buf, bytes, sink, 64);
        run<COOP4>("coop4x32", buf, bytes, sink, 128);
        run<COOP8>("coop8x16", buf, bytes, sink, 128);
        run<COOP32>("coop32x4", buf, bytes, sink, 128);
        run_tma<128>("tma128", buf, bytes, sink);
        run_tma<32>("tma32", buf, bytes, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    for (int c : {2, 3, 4, 6, 8}) {
        run_depth<1>(buf, 1840ull << 20, sink, c);
        run_depth<2>(buf, 1840ull << 20, sink, c);
        run_depth<4>(buf, 1840ull << 20, sink, c);
        run_depth<8>(buf, 1840ull << 20, sink, c);
    }
    cudaDeviceSynchronize();
    return 0;
}
