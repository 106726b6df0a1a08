// Does programmatic dependent launch let the CTAs of launch k+1 start on SMs that launch k has already left, while the
// slowest CTAs of launch k are still running (one CTA per SM, ~200 KB of shared memory each — the rollout kernel's shape)?
// Every CTA triggers launch_dependents at its start and never executes griddepcontrol.wait; CTA durations are staggered.
// Prints, per launch: first/last CTA start, first/last CTA end (us, relative to the first start of launch 0).
// Variants: plain launches; PDL; PDL with an event record between launches; PDL with a tiny kernel between launches.
//   nvcc -arch=sm_100a -O3 -o pdl_chain pdl_chain.cu && ./pdl_chain
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ unsigned long long g_t[8][160][2];

__device__ __forceinline__ unsigned long long now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void work(int launch, int base_us, int spread_us)
{
    extern __shared__ unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;");
    const unsigned long long t0 = now();
    if (threadIdx.x == 0) g_t[launch][blockIdx.x][0] = t0;
    smem[threadIdx.x] = (unsigned char)launch;
    // CTA b runs base + spread * (b mod 8) / 8 microseconds
    const unsigned long long dur = 1000ull * (unsigned long long)(base_us + spread_us * (int)(blockIdx.x % 8) / 8);
    while (now() - t0 < dur) { }
    if (threadIdx.x == 0) g_t[launch][blockIdx.x][1] = now();
}

__global__ void tiny(int *p) { if (threadIdx.x == 0 && p) *p = 1; }

static void run(const char *name, bool pdl, int between, int n_launch, int grid)
{
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(work, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(nullptr, 0, 0);
    for (int rep = 0; rep < 2; ++rep) {            // first repetition warms up
        for (int k = 0; k < n_launch; ++k) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = (pdl && k > 0) ? 1 : 0;
            cudaLaunchKernelEx(&cfg, work, k, 200, 100);
            if (between == 1) cudaEventRecord(ev, st);
            if (between == 2) tiny<<<1, 32, 0, st>>>(nullptr);
        }
        cudaStreamSynchronize(st);
    }
    cudaError_t e = cudaGetLastError();
    static unsigned long long h[8][160][2];
    cudaMemcpyFromSymbol(h, g_t, sizeof h);
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < grid; ++b) t0 = h[0][b][0] < t0 ? h[0][b][0] : t0;
    printf("== %s (%s)\n", name, cudaGetErrorString(e));
    for (int k = 0; k < n_launch; ++k) {
        unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
        for (int b = 0; b < grid; ++b) {
            s0 = h[k][b][0] < s0 ? h[k][b][0] : s0; s1 = h[k][b][0] > s1 ? h[k][b][0] : s1;
            e0 = h[k][b][1] < e0 ? h[k][b][1] : e0; e1 = h[k][b][1] > e1 ? h[k][b][1] : e1;
        }
        printf("  launch %d: CTA starts %.1f .. %.1f us, ends %.1f .. %.1f us\n", k, (s0 - t0) / 1e3, (s1 - t0) / 1e3,
               (e0 - t0) / 1e3, (e1 - t0) / 1e3);
    }
    cudaStreamDestroy(st);
}

int main()
{
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", n_sm);
    run("plain launches", false, 0, 4, n_sm);
    run("PDL, nothing between", true, 0, 4, n_sm);
    run("PDL, event record between", true, 1, 4, n_sm);
    run("PDL, tiny kernel between", true, 2, 4, n_sm);
    return 0;
}
