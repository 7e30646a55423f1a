// Bulk asynchronous copies (cp.async.bulk = the TMA unit, SASS UBLKCP, completion on an mbarrier) for the 4 KB rows the
// Griffin-Lim iteration kernel fetches per frame: 8 warps per CTA, each with two mbarriers; lane 0 copies a 4112-byte row
// global -> shared, all lanes wait and check it.  All modes pass (profiles/r2/bulk_copy_micro.txt).
// Round-2 experiment behind it (profiles/r2/sweep_bulk.txt): k_gl_stream with its magnitude rows fetched this way instead of
// nine cp.async per lane ran 16.64 instead of 16.24 ms per 60-iteration launch (one lane's issue + the mbarrier round trip
// are exposed where 32 lanes x LDGSTS overlapped with the FFT), and the variant that also staged the next frame's samples
// this way hung; the kernel keeps cp.async.
#include <cstdio>
#include <cuda_runtime.h>
namespace nsb {
// ---- bulk asynchronous copies (the TMA unit: SASS UBLKCP) completing on an mbarrier --------------------------------------
// One elected lane moves a whole 4 KB row global -> shared with ONE instruction; the other 31 lanes issue nothing (the
// cp.async form costs every lane nine LDGSTS plus their address arithmetic per row).  The data comes from L2 (the TMA unit
// does not allocate in L1): safe for the waveform that other SMs rewrite inside the launch.
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// (the calling lane) order this warp's earlier generic-proxy accesses of the destination before the async-proxy write, post the
// byte count, start the copy; dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;\n"
                 "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];"
                 ::"r"(b), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(b), "r"(parity) : "memory");
}
}  // namespace nsb
using namespace nsb;

__global__ void __launch_bounds__(256, 2) k_bulk(const float* src, int rows, int iters, int* errors, int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* mbar_all = reinterpret_cast<unsigned long long*>(smem);
    float* tiles = reinterpret_cast<float*>(smem + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = tiles + warp * 2112;
    unsigned long long* mb = mbar_all + 2 * warp;
    unsigned ph = 0, ph2 = 0;
    if (threadIdx.x < 16) mbar_init(mbar_all + threadIdx.x, 1);
    mbar_fence_init();
    __syncthreads();
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        const int row = (blockIdx.x * 8 + warp + it * 37) % rows;
        const float* s = src + (size_t)row * 1056;
        if (mode & 2) { if (lane == 0) bulk_prefetch_l2(src + (size_t)((row + 1) % rows) * 1056, 4112); }
        __syncwarp();
        if (lane == 0) bulk_load(tile, s, 4112, mb);
        __syncwarp();
        mbar_wait(mb, ph); ph ^= 1;
        for (int i = lane; i < 1028; i += 32) bad += (tile[i] != s[i]);
        __syncwarp();
        if (mode & 1) {           // second barrier, offset destination (the staged samples)
            if (lane == 0) bulk_load(tile + 512, s + 4, 4112, mb + 1);
            mbar_wait(mb + 1, ph2); ph2 ^= 1;
            for (int i = lane; i < 1028; i += 32) bad += (tile[512 + i] != s[4 + i]);
            __syncwarp();
        }
        if (mode & 4) __syncthreads();
    }
    if (bad) atomicAdd(errors, bad);
}

int main() {
    const int rows = 4096;
    float* d; int* e;
    cudaMalloc(&d, sizeof(float) * 1056 * (rows + 1));
    cudaMalloc(&e, 4);
    float* h = new float[1056 * (rows + 1)];
    for (int i = 0; i < 1056 * (rows + 1); ++i) h[i] = (float)(i % 9973);
    cudaMemcpy(d, h, sizeof(float) * 1056 * (rows + 1), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 8 * 8448);
    for (int mode = 0; mode < 8; ++mode) {
        cudaMemset(e, 0, 4);
        k_bulk<<<296, 256, 128 + 8 * 8448>>>(d, rows, 200, e, mode);
        cudaError_t err = cudaDeviceSynchronize();
        int errs = -1;
        cudaMemcpy(&errs, e, 4, cudaMemcpyDeviceToHost);
        printf("mode %d: %s, mismatches %d\n", mode, cudaGetErrorString(err), errs);
        fflush(stdout);
        if (err != cudaSuccess) return 1;
    }
    return 0;
}
