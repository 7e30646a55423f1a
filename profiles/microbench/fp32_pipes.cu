// Microbenchmark: per-SM issue rates of the instructions the FFT kernels are made of (B200, sm_100a).
// Prints lane-ops per clock per SM for independent chains at 16 and 32 resident warps per SM.
// Used to set the FP32 roofline denominator honestly (is 3-register FFMA full rate? does f32x2 help?).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void k(float* out, long long* cycles, float seed) {
    float a[CHAINS], b[CHAINS];
    unsigned long long pa[CHAINS / 2];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = 1.0f + i * 1e-4f; }
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) pa[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
    const float c0 = seed * 0.5f, c1 = seed * 0.25f;
    unsigned long long pb = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c1);
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
            if (OP == 1) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(a[i]) : "f"(c0));      // immediate multiplier
            if (OP == 4 && i < CHAINS / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pa[i]) : "l"(pb));
            if (OP == 5 && i < CHAINS / 2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb));
            if (OP == 6) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 7) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
            if (OP == 8) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + i * 32) & 4095)))); a[i] += v; }
            if (OP == 9) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + ((2 * threadIdx.x + i * 64) & 4094)))); a[i] += v.x + v.y; }
            if (OP == 10) asm volatile("add.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(a[(i + 1) % CHAINS]), "f"(b[i]));   // FADD, 2 distinct sources
            if (OP == 11) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[(i + 1) % CHAINS]), "f"(b[i]), "f"(b[(i + 3) % CHAINS])); // 3 distinct
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) s += __uint_as_float((unsigned)pa[i]) + __uint_as_float((unsigned)(pa[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_thread_ops, int flops_per_op) {
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {512, 1024}) {
        float* out; long long* cyc;
        cudaMalloc(&out, sizeof(float) * dev_sms * threads);
        cudaMalloc(&cyc, sizeof(long long) * dev_sms);
        k<OP><<<dev_sms, threads>>>(out, cyc, 1.0f);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<OP><<<dev_sms, threads>>>(out, cyc, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * dev_sms, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < dev_sms; ++i) avg += h[i]; avg /= dev_sms;
        double ops = (double)ITERS * per_thread_ops * threads;   // lane-ops per SM
        printf("%-28s warps/SM=%2d  lane-ops/clk/SM=%7.1f  (%.1f values/clk/SM)  clk=%.0f  ms=%.3f  eff.GHz=%.3f\n", name, threads / 32,
               ops / avg, ops * flops_per_op / avg, avg, ms, avg / (ms * 1e6));
        cudaFree(out); cudaFree(cyc);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device: %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("FADD r,r,r(same chain)", CHAINS, 1);
    run<1>("FMUL", CHAINS, 1);
    run<2>("FFMA d=d*b+c", CHAINS, 1);
    run<3>("FFMA imm", CHAINS, 1);
    run<4>("FFMA2 (f32x2)", CHAINS / 2, 2);
    run<5>("FADD2 (f32x2)", CHAINS / 2, 2);
    run<6>("MUFU.RSQ", CHAINS, 1);
    run<7>("SHFL.BFLY", CHAINS, 1);
    run<8>("LDS.32 (+FADD)", CHAINS, 1);
    run<9>("LDS.64 (+2 FADD)", CHAINS, 2);
    run<10>("FADD 2 distinct srcs", CHAINS, 1);
    run<11>("FFMA 3 distinct srcs", CHAINS, 1);
    return 0;
}
