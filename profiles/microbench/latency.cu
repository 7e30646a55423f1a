// Microbenchmark: dependent-issue latency of the instructions the FFT kernels are made of, and the throughput a
// scheduler reaches with W warps per SM sub-partition each running C independent chains (how much ILP does a warp
// need at 4 warps per scheduler?).  B200, sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int OP, int C>
__global__ void k(float* out, long long* cycles, float seed) {
    float a[C];
    unsigned long long pa[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { a[i] = OP == 6 ? 0.f : seed + threadIdx.x * 1e-3f + i; pa[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float c0 = seed * 0.5f, c1 = seed * 0.25f;
    unsigned long long pb = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c1);
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    unsigned sbase = (unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 31));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c0));
            if (OP == 2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb));
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pa[i]) : "l"(pb));
            if (OP == 4) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb));
            if (OP == 5) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 6) { unsigned off; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(off) : "r"(sbase + (unsigned)__float_as_uint(a[i]))); a[i] = __uint_as_float(off); }   // pointer chase (table is zeros)
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i] + __uint_as_float((unsigned)pa[i]) + __uint_as_float((unsigned)(pa[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int C>
void run(const char* name, int warps_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    for (int r = 0; r < 2; ++r) k<OP, C><<<sms, warps_per_sm * 32>>>(out, cyc, OP == 6 ? 0.f : 1.0f);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    double per_sched = (double)ITERS * C * (warps_per_sm < 4 ? 1 : warps_per_sm / 4);   // warp-instructions per scheduler
    printf("%-10s warps/scheduler=%d chains/warp=%d  cycles per chain step = %6.2f   issue rate per scheduler = %.3f inst/clk\n", name,
           warps_per_sm < 4 ? 1 : warps_per_sm / 4, C, avg / ITERS, per_sched / avg);
    cudaFree(out); cudaFree(cyc);
}

template <int OP>
void all(const char* name) {
    run<OP, 1>(name, 1);
    run<OP, 1>(name, 16); run<OP, 2>(name, 16); run<OP, 4>(name, 16); run<OP, 8>(name, 16);
    run<OP, 1>(name, 24); run<OP, 2>(name, 24);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device: %s, %d SMs\n", p.name, p.multiProcessorCount);
    all<0>("FADD"); all<1>("FFMA"); all<2>("FADD2"); all<3>("FFMA2"); all<4>("FMUL2"); all<5>("MUFU.RSQ"); all<6>("LDS chase");
    return 0;
}
