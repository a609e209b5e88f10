// Micro-benchmark: how many DRAM bytes does one random 32-byte probe cost on B200, per load flavour?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_variants gather_variants.cu
//   ./gather_variants [log2_bytes=32]          (run under ncu --metrics dram__bytes_read.sum for the traffic)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; return x; }
template <int V> __device__ __forceinline__ u64 load32(const u64 *p) {
    u64 a, b, c, d;
    if(V == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if(V == 1) { asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
                 asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 2)); }
    if(V == 2) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if(V == 3) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if(V == 4) asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if(V == 5) asm volatile("ld.global.lu.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if(V == 6) { asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p)); b = c = d = 0; }   // 8 bytes only
    if(V == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}
template <int V> __global__ void gather(const u64 *t, unsigned bits, u64 n, u64 seed, u64 *out) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nt = (u64)gridDim.x * blockDim.x;
    u64 acc = 0;
    for(u64 i = tid * 4; i < n; i += nt * 4) {
        u64 v[4];
#pragma unroll
        for(int j = 0; j < 4; ++j) v[j] = load32<V>(t + ((mix64(seed + i + j) >> (64 - bits)) << 2));
        acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    if(acc == 0x1234567) *out = acc;
}
template <int V> void run(const char *name, const u64 *t, unsigned bits, u64 n, u64 *out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<V><<<148 * 8, 256>>>(t, bits, n / 8, 1, out);
    cudaEventRecord(a);
    gather<V><<<148 * 8, 256>>>(t, bits, n, 7, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %8.3f ms  %7.1f G probes/s  %7.1f GB/s of 32-byte sectors\n", name, ms, n / ms / 1e6, n * 32.0 / ms / 1e6);
}
int main(int argc, char **argv) {
    const unsigned lg = argc > 1 ? atoi(argv[1]) : 32;            // table bytes = 2^lg
    const int gran = argc > 2 ? atoi(argv[2]) : -1;
    if(gran >= 0) printf("cudaDeviceSetLimit(MaxL2FetchGranularity, %d) -> %d\n", gran, (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu, table 2^%u bytes\n", g, lg);
    u64 *t, *out; cudaMalloc(&t, 1ull << lg); cudaMalloc(&out, 8); cudaMemset(t, 0xff, 1ull << lg);
    const unsigned bits = lg - 5; const u64 n = 1ull << 28;
    run<0>("ld.nc.L1::no_allocate.v4.u64", t, bits, n, out);
    run<1>("2 x ld.nc.L1::no_allocate.v2.u64", t, bits, n, out);
    run<2>("ld.global.v4.u64 (ca)", t, bits, n, out);
    run<3>("ld.global.cg.v4.u64", t, bits, n, out);
    run<4>("ld.global.cs.v4.u64", t, bits, n, out);
    run<5>("ld.global.lu.v4.u64", t, bits, n, out);
    run<6>("ld.nc.u64 (8 bytes)", t, bits, n, out);
    run<7>("ld.nc.L2::evict_first.v4.u64", t, bits, n, out);
    return 0;
}
