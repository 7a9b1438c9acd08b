// Microbenchmark 2: random CHUNK reads (chunk = 32..4096 contiguous bytes, read by a group of
// lanes) and concurrency sweep for random 32 B sectors, on a 16 GB table.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
// LANES lanes cooperate on one chunk of LANES*32 bytes
template <int LANES>
__global__ void probe_chunk(const uint8_t* __restrict__ tbl, uint64_t nchunks, int64_t iters, uint64_t* sink, uint64_t seed) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t grp = tid / LANES; int sub = tid % LANES;
    uint64_t acc = 0;
    for (int64_t it = 0; it < iters; it++) {
        uint64_t r = mix(seed + grp * 0x10001ULL + (uint64_t)it * 0x9E3779B97F4A7C15ULL);
        const uint8_t* p = tbl + __umul64hi(r, nchunks) * (LANES * 32ULL) + sub * 32;
        uint64_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        acc ^= a ^ b ^ c ^ d;
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <int LANES>
void run_chunk(const uint8_t* tbl, uint64_t bytes, uint64_t* sink, int threads, int bps, const char* tag) {
    int64_t iters = 256; uint64_t nchunks = bytes / (LANES * 32ULL);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a)); probe_chunk<LANES><<<148 * bps, threads>>>(tbl, nchunks, iters, sink, 5 + rep);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    double n = 148.0 * bps * threads * iters;
    printf("%s chunk %5d B  thr/SM %4d : %7.2f G chunks/s %7.2f G sectors/s %7.1f GB/s\n", tag, LANES * 32, threads * bps, n / LANES / best / 1e6, n / best / 1e6, n * 32 / best / 1e6);
    fflush(stdout);
}
int main(int argc, char** argv) {
    // argv[1] = GiB to ALLOCATE (the probes always use the first 16 GiB): does the size of the
    // allocation change the random-access rate (page size / TLB reach)?
    uint64_t alloc = (argc > 1 ? atoll(argv[1]) : 16) * (1ULL << 30);
    uint64_t bytes = 16ULL << 30; uint8_t* tbl; uint64_t* sink;
    if (alloc < bytes) alloc = bytes;
    CK(cudaMalloc(&tbl, alloc)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(tbl, 1, alloc));
    printf("allocated %.1f GiB, probing the first 16 GiB\n", alloc / 1073741824.0);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", g);
    int thr_cfg[][2] = {{64, 1}, {128, 1}, {256, 1}, {256, 2}, {256, 3}, {256, 4}, {256, 6}, {256, 8}};
    for (auto& c : thr_cfg) run_chunk<1>(tbl, bytes, sink, c[0], c[1], "gran-default");
    run_chunk<2>(tbl, bytes, sink, 256, 8, "gran-default");
    run_chunk<4>(tbl, bytes, sink, 256, 8, "gran-default");
    run_chunk<8>(tbl, bytes, sink, 256, 8, "gran-default");
    run_chunk<16>(tbl, bytes, sink, 256, 8, "gran-default");
    run_chunk<32>(tbl, bytes, sink, 256, 8, "gran-default");
    for (size_t gr : {32, 128}) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gr);
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("set L2 fetch granularity %zu -> %s, now %zu\n", gr, cudaGetErrorString(e), g);
        char tag[32]; snprintf(tag, 32, "gran-%zu", gr);
        run_chunk<1>(tbl, bytes, sink, 256, 4, tag);
        run_chunk<1>(tbl, bytes, sink, 256, 8, tag);
        run_chunk<4>(tbl, bytes, sink, 256, 8, tag);
    }
    return 0;
}
