// Latency of a warp-wide dependent gather of u16 cells (the MVA grid lookup pattern): cycles per load for one warp on
// an otherwise idle SM, as a function of how the 32 lane addresses spread.  nvcc -arch=sm_100a -O3 -o gather gather_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void chase(const uint16_t *g, uint32_t n, int mode, int iters, long long *out, uint32_t *sink)
{
    uint32_t lane = threadIdx.x & 31;
    uint32_t x = (blockIdx.x * 2654435761u) ^ (lane * 40503u + 12345u);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u + acc;            // depends on the previous load
        uint32_t idx;
        if (mode == 0) idx = x % n;                                  // 32 random lines of the whole array
        else if (mode == 1) idx = ((x >> 5 << 5) % (n - 64)) + lane;   // the warp's lanes in one 64-byte run (random place)
        else if (mode == 2) idx = x % 8192;                          // 32 random lines of a 16 KB (L1-resident) region
        else idx = (__shfl_sync(0xFFFFFFFFu, x, 0) % (n - 64));      // all lanes the same address
        acc = __ldg(g + idx) & 1u;
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (t1 - t0) / iters;
    if (x == 12345u) *sink = x + acc;
}

int main()
{
    const uint32_t n = 1284 * 1030;           // the LOWW grid at 0.0625 nm
    uint16_t *g; long long *out; uint32_t *sink;
    cudaMalloc(&g, n * 2); cudaMemset(g, 0, n * 2);
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
    const char *names[4] = {"32 random lines, 2.6 MB", "one 64-byte run, random place", "32 random lines, 16 KB region", "one address"};
    for (int ctas = 1; ctas <= 148; ctas *= 148) {
        for (int mode = 0; mode < 4; ++mode) {
            chase<<<ctas, 32>>>(g, n, mode, 2000, out, sink);      // warm-up: brings the array into L2
            chase<<<ctas, 32>>>(g, n, mode, 4000, out, sink);
            long long h[148];
            cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
            long long mn = h[0], mx = h[0];
            for (int i = 1; i < ctas; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
            printf("CTAs %3d  %-32s  %lld .. %lld cycles per dependent load (incl. ~25 of index arithmetic)\n", ctas, names[mode], mn, mx);
        }
    }
    return 0;
}
