// Microbenchmark: instruction supply.  A loop body of N KB of straight-line FFMA code, run by 16 warps per SM that are
// either in step or spread over the body (each warp starts the loop after its own delay).  Throughput against body size
// shows where the instruction caches stop covering a kernel whose warps are NOT in lockstep - the situation of the
// Griffin-Lim iteration kernels (40 KB of straight-line code per frame).  B200, sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

#define F8 asm volatile("fma.rn.f32 %0, %0, %8, %9;\n fma.rn.f32 %1, %1, %8, %9;\n fma.rn.f32 %2, %2, %8, %9;\n fma.rn.f32 %3, %3, %8, %9;\n" \
                        "fma.rn.f32 %4, %4, %8, %9;\n fma.rn.f32 %5, %5, %8, %9;\n fma.rn.f32 %6, %6, %8, %9;\n fma.rn.f32 %7, %7, %8, %9;" \
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));
#define F64 F8 F8 F8 F8 F8 F8 F8 F8          /* 64 instructions = 1 KB of SASS */
#define K4 F64 F64 F64 F64
#define K16 K4 K4 K4 K4

template <int KB> struct Body;
#define BODY(kb, code) template <> struct Body<kb> { static __device__ __forceinline__ void run(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5, float& a6, float& a7, float b, float c) { code } };
BODY(4, K4)
BODY(8, K4 K4)
BODY(16, K16)
BODY(24, K16 K4 K4)
BODY(32, K16 K16)
BODY(48, K16 K16 K16)
BODY(64, K16 K16 K16 K16)
BODY(96, K16 K16 K16 K16 K16 K16)
BODY(128, K16 K16 K16 K16 K16 K16 K16 K16)
BODY(192, K16 K16 K16 K16 K16 K16 K16 K16 K16 K16 K16 K16)

template <int KB>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int reps, int spread, float b, float c) {
    float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    long long t0 = clock64();
    if (spread) {                         // warp w enters the loop w * (body time / 16) later: 16 different places in the body
        const long long wait = (long long)warp * KB * 64 * 4 / 16;   // ~4 cycles per instruction at 4 warps per scheduler
        while (clock64() - t0 < wait) {}
    }
    for (int r = 0; r < reps; ++r) Body<KB>::run(a0, a1, a2, a3, a4, a5, a6, a7, b, c);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KB>
void run() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int reps = 4096 / KB;            // same instruction count for every body size
    for (int spread = 0; spread < 2; ++spread) {
        float ms = 0;
        for (int it = 0; it < 2; ++it) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k<KB><<<sms, 512>>>(out, cyc, reps, spread, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double inst = (double)reps * KB * 64 * 16;      // warp-instructions per SM
        printf("body %3d KB  warps %-9s  %.3f ms  %.2f warp-instructions/clk/SM (peak 4)\n", KB, spread ? "spread" : "in step", ms,
               inst / (ms * 1e-3 * 1.965e9));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<4>(); run<8>(); run<16>(); run<24>(); run<32>(); run<48>(); run<64>(); run<96>(); run<128>(); run<192>();
    return 0;
}
