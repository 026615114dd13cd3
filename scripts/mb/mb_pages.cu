// Micro-benchmark (measurement tool, not product): is the random-gather cliff past ~256 MB a matter of how many
// distinct 2 MB pages are touched (address translation) or of DRAM?  Random 32-byte-sector and 128-byte-line gathers
//   (1) over P pages of 2 MB spread with a stride inside a large cudaMalloc buffer (footprint P x 2 MB, span P x stride x 2 MB),
//   (2) over a buffer mapped with the virtual-memory API from ONE physical handle, VA aligned to 1 GB
//       (whether the driver then uses a larger page size shows as the cliff moving),
//   (3) over managed memory prefetched to the device.
// Output: CSV  alloc,pages,stride,footprint_MB,width_B,G_access_per_s,GB_per_s
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld256(const uint32_t *p) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

constexpr uint64_t PAGE = 2ull << 20;

template <int WIDTH, int UNROLL>
__global__ void gather(const uint8_t *__restrict__ buf, uint32_t pages, uint32_t stride, int iters, uint32_t *sink) {
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            const uint32_t pg = (uint32_t)(((x >> 32) * (uint64_t)pages) >> 32);
            const uint32_t line = (uint32_t)x & (uint32_t)(PAGE / 128 - 1);
            const uint8_t *p = buf + (uint64_t)pg * stride * PAGE + (uint64_t)line * 128;
            if (WIDTH == 32) {
                acc ^= ld256((const uint32_t *)(p + ((x >> 20) & 3) * 32));
            } else {
#pragma unroll
                for (int w = 0; w < 4; ++w) acc ^= ld256((const uint32_t *)(p + w * 32));
            }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int WIDTH>
static double run(const uint8_t *buf, uint32_t pages, uint32_t stride, uint32_t *sink) {
    const int grid = 148 * 4, block = 256, UN = 10;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 40;
    gather<WIDTH, UN><<<grid, block>>>(buf, pages, stride, 20, sink);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        gather<WIDTH, UN><<<grid, block>>>(buf, pages, stride, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < 10.f && rep == 0 && iters < 640) { iters *= 4; rep = -1; continue; }
        const double rate = (double)grid * block * iters * UN / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

static void line(const char *alloc, const uint8_t *buf, uint32_t pages, uint32_t stride, uint32_t *sink) {
    const double r32 = run<32>(buf, pages, stride, sink), r128 = run<128>(buf, pages, stride, sink);
    const double mb = pages * 2.0;
    printf("%s,%u,%u,%.0f,32,%.2f,%.1f\n", alloc, pages, stride, mb, r32 / 1e9, r32 * 32 / 1e9);
    printf("%s,%u,%u,%.0f,128,%.2f,%.1f\n", alloc, pages, stride, mb, r128 / 1e9, r128 * 128 / 1e9);
    fflush(stdout);
}

#define DRV(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_ = nullptr; cuGetErrorString(r_, &s_); \
    fprintf(stderr, "%s -> %s\n", #x, s_ ? s_ : "?"); ok = false; } } while (0)

int main(int argc, char **argv) {
    const bool quick = argc > 1;  // any argument: the short list (used under ncu)
    cudaFree(0);
    uint32_t *sink = nullptr;
    cudaMalloc(&sink, 4);
    printf("alloc,pages,stride,footprint_MB,width_B,G_access_per_s,GB_per_s\n");

    // (1) plain cudaMalloc, 16 GB
    const uint64_t big = 16ull << 30;
    uint8_t *buf = nullptr;
    if (cudaMalloc(&buf, big) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemset(buf, 0x5a, big);
    fprintf(stderr, "cudaMalloc base %p (mod 512 MB = %llu MB)\n", (void *)buf, (unsigned long long)(((uintptr_t)buf) & ((512ull << 20) - 1)) >> 20);
    if (quick) {
        line("cudaMalloc", buf, 8192, 1, sink);
        line("cudaMalloc", buf, 100, 1, sink);
        cudaDeviceSynchronize();
        return 0;
    }
    for (uint32_t p : {64u, 100u, 115u, 128u, 144u, 160u, 192u, 230u, 256u, 512u, 2048u, 8192u}) line("cudaMalloc", buf, p, 1, sink);
    for (uint32_t p : {64u, 100u, 115u, 128u, 160u, 256u}) line("cudaMalloc", buf, p, 16, sink);   // same footprints spread over 16x the span
    for (uint32_t p : {100u, 128u}) line("cudaMalloc", buf, p, 64, sink);
    cudaFree(buf);

    // (2) virtual-memory API: one physical handle, VA aligned to 1 GB
    {
        bool ok = true;
        CUdevice dev;
        DRV(cuDeviceGet(&dev, 0));
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = 0;
        size_t gmin = 0, grec = 0;
        DRV(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
        DRV(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        fprintf(stderr, "allocation granularity: minimum %zu, recommended %zu\n", gmin, grec);
        const size_t sz = 16ull << 30;
        CUmemGenericAllocationHandle h = 0;
        CUdeviceptr va = 0;
        DRV(cuMemCreate(&h, sz, &prop, 0));
        if (ok) DRV(cuMemAddressReserve(&va, sz, 1ull << 30, 0, 0));
        if (ok) DRV(cuMemMap(va, sz, 0, h, 0));
        CUmemAccessDesc acc = {};
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (ok) DRV(cuMemSetAccess(va, sz, &acc, 1));
        if (ok) {
            cudaMemset((void *)va, 0x5a, sz);
            fprintf(stderr, "vmm base %p\n", (void *)va);
            for (uint32_t p : {100u, 128u, 160u, 230u, 256u, 512u, 2048u, 8192u}) line("vmm_1handle_1GBaligned", (const uint8_t *)va, p, 1, sink);
            line("vmm_1handle_1GBaligned", (const uint8_t *)va, 128, 16, sink);
            cuMemUnmap(va, sz);
            cuMemAddressFree(va, sz);
            cuMemRelease(h);
        }
    }
    // (3) managed memory, preferred location + prefetch to the device
    {
        uint8_t *m = nullptr;
        const size_t sz = 8ull << 30;
        if (cudaMallocManaged(&m, sz) == cudaSuccess) {
            cudaMemLocation loc = {};
            loc.type = cudaMemLocationTypeDevice;
            loc.id = 0;
            cudaMemAdvise(m, sz, cudaMemAdviseSetPreferredLocation, loc);
            cudaMemPrefetchAsync(m, sz, loc, 0, 0);
            cudaMemset(m, 0x5a, sz);
            cudaDeviceSynchronize();
            for (uint32_t p : {100u, 230u, 512u, 4096u}) line("managed_prefetched", m, p, 1, sink);
            cudaFree(m);
        } else fprintf(stderr, "cudaMallocManaged failed\n");
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
