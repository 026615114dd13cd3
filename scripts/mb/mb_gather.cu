// Micro-benchmark (measurement tool, not product): random-gather throughput of one B200 as a function of
// table footprint and access width.  Every lane keeps UNROLL independent loads in flight.
//   width 4   : one 4-byte load per access (what a single bloom-bit probe does)
//   width 32  : one 256-bit load  (one 32-byte sector: a 256-node sliced row)
//   width 64  : two 256-bit loads to adjacent sectors
//   width 128 : four 256-bit loads (one full L2 line)
// Output: CSV  footprint_MB,width,G_access_per_s,GB_per_s
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld256(const uint32_t *p) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int WIDTH, int UNROLL>
__global__ void gather(const uint32_t *__restrict__ buf, uint64_t n_units, int iters, uint32_t *sink) {
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            const uint64_t u = (uint64_t)(((unsigned __int128)x * n_units) >> 64);
            if (WIDTH == 4) {
                acc += __ldg(buf + u * 8 + (x & 7));
            } else {
                const uint32_t *p = buf + u * (WIDTH / 4);
#pragma unroll
                for (int w = 0; w < WIDTH / 32; ++w) acc ^= ld256(p + w * 8);
            }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int WIDTH>
static double run(const uint32_t *buf, uint64_t bytes, uint32_t *sink, int ctas_per_sm) {
    const uint64_t unit = WIDTH == 4 ? 32 : WIDTH;
    const uint64_t n_units = bytes / unit;
    const int grid = 148 * ctas_per_sm, block = 256, UN = 10;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 40;
    gather<WIDTH, UN><<<grid, block>>>(buf, n_units, 20, sink);  // warm the cache / TLB
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        gather<WIDTH, UN><<<grid, block>>>(buf, n_units, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < 20.f && rep == 0) { iters *= 4; rep = -1; continue; }
        const double rate = (double)grid * block * iters * UN / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int main(int argc, char **argv) {
    std::vector<double> mbs = {8, 16, 32, 48, 57.5, 64, 80, 96, 115, 128, 160, 230, 460, 920, 2048, 8192, 36864};
    if (argc > 1) { mbs.clear(); for (int i = 1; i < argc; ++i) mbs.push_back(atof(argv[i])); }
    double maxmb = 0;
    for (double m : mbs) if (m > maxmb) maxmb = m;
    if (const char *g = getenv("MB_L2_FETCH")) {  // cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "L2 fetch granularity: asked %s, %s, now %zu\n", g, cudaGetErrorString(e), got);
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "L2 fetch granularity (default): %zu\n", got);
    }
    uint32_t *buf = nullptr, *sink = nullptr;
    const uint64_t cap = (uint64_t)(maxmb * 1048576.0) + 4096;
    if (cudaMalloc(&buf, cap) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
    cudaMemset(buf, 0x5a, cap);
    printf("footprint_MB,width_B,ctas_per_sm,G_access_per_s,GB_per_s\n");
    for (double m : mbs) {
        const uint64_t bytes = ((uint64_t)(m * 1048576.0)) & ~(uint64_t)127;
        for (int c : {4, 8}) {
            double r4 = run<4>(buf, bytes, sink, c), r32 = run<32>(buf, bytes, sink, c), r64 = run<64>(buf, bytes, sink, c),
                   r128 = run<128>(buf, bytes, sink, c);
            printf("%.1f,4,%d,%.2f,%.1f\n", m, c, r4 / 1e9, r4 * 32 / 1e9);
            printf("%.1f,32,%d,%.2f,%.1f\n", m, c, r32 / 1e9, r32 * 32 / 1e9);
            printf("%.1f,64,%d,%.2f,%.1f\n", m, c, r64 / 1e9, r64 * 64 / 1e9);
            printf("%.1f,128,%d,%.2f,%.1f\n", m, c, r128 / 1e9, r128 * 128 / 1e9);
            fflush(stdout);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
