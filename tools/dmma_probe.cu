// Microbenchmark: FP64 DMMA (mma.sync.m8n8k4.f64) throughput vs warps/SM and independent accumulators per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NACC, bool MUL>
__global__ void probe(double* out, int iters, double seed) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed, s = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            double av = a;
            if (MUL) { av = a * s; s += 1e-12; }   // a dependent DMUL feeding each DMMA
            dmma(acc[i], av, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += acc[i][0] + acc[i][1];
    if (r == 12345.678) out[0] = r;
}
template <int NACC, bool MUL>
void run(int blocks_per_sm, int warps, int sms, double* d) {
    int iters = 40000 / NACC * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<NACC, MUL><<<sms * blocks_per_sm, warps * 32>>>(d, iters, 0.25);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = (double)sms * blocks_per_sm * warps * iters * NACC * 512.0;
    printf("acc=%2d mul=%d blocks/SM=%d warps/blk=%d (warps/SMSP=%.1f): %6.2f TFLOP/s\n", NACC, (int)MUL, blocks_per_sm, warps,
           blocks_per_sm * warps / 4.0, fl / (best * 1e-3) / 1e12);
}
// which hardware warp slot (%warpid; slot % 4 = SM sub-partition) the warps of co-resident CTAs get
__global__ void slot_map(int* out, int warps) {
    extern __shared__ double pad[];
    unsigned sm, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        int* o = out + 3 * (blockIdx.x * warps + (threadIdx.x >> 5));
        o[0] = sm; o[1] = wid; o[2] = blockIdx.x;
    }
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {}
}
static void report_slots(int sms, int warps, int regs_smem_kb) {
    int* d; int n = sms * 2 * warps;
    cudaMalloc(&d, n * 3 * sizeof(int));
    cudaFuncSetAttribute(slot_map, cudaFuncAttributeMaxDynamicSharedMemorySize, regs_smem_kb * 1024);
    slot_map<<<sms * 2, warps * 32, regs_smem_kb * 1024>>>(d, warps);
    int* h = new int[n * 3];
    cudaMemcpy(h, d, n * 3 * sizeof(int), cudaMemcpyDeviceToHost);
    int sm0 = h[0];
    printf("warp slots on SM %d (2 CTAs x %d warps, %d KB smem each):", sm0, warps, regs_smem_kb);
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) if (h[3 * i] == sm0) { printf(" cta%d:%d", h[3 * i + 2], h[3 * i + 1]); cnt[h[3 * i + 1] & 3]++; }
    printf("\n  warps per sub-partition: %d %d %d %d\n", cnt[0], cnt[1], cnt[2], cnt[3]);
    cudaFree(d); delete[] h;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 8);
    report_slots(sms, 6, 93);
    report_slots(sms, 4, 60);
    // K3's shape (6 warps x 2 CTAs: 4/4/2/2 warps per sub-partition) against balanced shapes with 12 warps per SM
    for (int w : {6, 12, 3}) for (int b : {1, 2, 4}) {
        if (w * b > 16) continue;
        run<9, false>(b, w, sms, d); run<9, true>(b, w, sms, d);
    }
    for (int w : {4, 8}) for (int b : {1, 2, 4}) {
        run<1, false>(b, w, sms, d); run<2, false>(b, w, sms, d); run<4, false>(b, w, sms, d);
        run<9, false>(b, w, sms, d); run<18, false>(b, w, sms, d); run<18, true>(b, w, sms, d);
    }
    return 0;
}
