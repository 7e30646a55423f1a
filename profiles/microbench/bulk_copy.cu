// Stand-alone check of the bulk-copy helpers of gl_iter.cuh (cp.async.bulk + mbarrier), the way k_gl_stream<..., BULK> uses them:
// 8 warps per CTA, each with two mbarriers; lane 0 copies a 4112-byte row global -> shared, all lanes wait and check it.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nspeech_b200/csrc/gl_iter.cuh"
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
