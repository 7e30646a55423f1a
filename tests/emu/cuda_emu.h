// TEST INFRASTRUCTURE ONLY - a tiny CUDA-on-CPU shim so that the *unmodified* kernel and C-ABI sources
// (nspeech_b200/csrc/*.cu, *.cuh) can be compiled with g++ -DNSB_EMULATE and executed in the GPU-less
// build container: one OS thread per CUDA thread, blocks run one after another, __syncthreads /
// __syncwarp are real barriers, warp shuffles go through an exchange buffer.  It exists to catch
// indexing, tiling and synchronisation mistakes (it also runs under -fsanitize=thread) before GPU time
// is spent.  The product library is never built this way and the product loader never looks for the
// emulated library; only tests/ load it.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include <algorithm>
#include <array>

// ---- vector types -----------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r; r.x = a; r.y = b; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline int4 make_int4(int a, int b, int c, int d) { int4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
static inline float __int_as_float(int v) { float r; memcpy(&r, &v, 4); return r; }

// ---- qualifiers -------------------------------------------------------------------------------
#define __global__ inline
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

// ---- execution context --------------------------------------------------------------------------
namespace nsb_emu {
struct BlockCtx {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<std::array<double, 32>> xchg;
};
extern thread_local BlockCtx* g_ctx;
extern thread_local uint3 g_tid, g_bid;
extern thread_local dim3 g_bdim, g_gdim;
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);
unsigned char* dyn_smem();
}  // namespace nsb_emu
#define threadIdx (nsb_emu::g_tid)
#define blockIdx (nsb_emu::g_bid)
#define blockDim (nsb_emu::g_bdim)
#define gridDim (nsb_emu::g_gdim)

static inline void __syncthreads() { nsb_emu::g_ctx->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { nsb_emu::g_ctx->warp_bar[nsb_emu::g_tid.x >> 5]->arrive_and_wait(); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    auto* c = nsb_emu::g_ctx;
    int w = nsb_emu::g_tid.x >> 5, l = nsb_emu::g_tid.x & 31;
    c->xchg[w][l] = (double)v;
    c->warp_bar[w]->arrive_and_wait();
    T r = l >= delta ? (T)c->xchg[w][l - delta] : v;
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
    auto* c = nsb_emu::g_ctx;
    int w = nsb_emu::g_tid.x >> 5, l = nsb_emu::g_tid.x & 31;
    c->xchg[w][l] = (double)v;
    c->warp_bar[w]->arrive_and_wait();
    T r = (T)c->xchg[w][l ^ mask];
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __saturatef(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
using std::isfinite;
using std::min;
using std::max;

// ---- runtime API (host side) --------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct CUstream_emu* cudaStream_t;
typedef struct CUevent_emu* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; size_t sharedMemPerBlockOptin; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 3; p->sharedMemPerBlockOptin = 232448; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(0x1); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(0x1); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
