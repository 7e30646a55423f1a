// TEST INFRASTRUCTURE ONLY: thread-per-CUDA-thread block executor for cuda_emu.h, plus lane-level
// entry points used by tests/test_emulated_kernels.py.  See cuda_emu.h.
#include <mutex>
#include "cuda_emu.h"
#include "../../nspeech_b200/csrc/frame_fft.cuh"

namespace nsb_emu {
thread_local BlockCtx* g_ctx = nullptr;
thread_local uint3 g_tid, g_bid;
thread_local dim3 g_bdim, g_gdim;
alignas(256) static unsigned char g_smem[256 * 1024];
unsigned char* dyn_smem() { return g_smem; }

// one kernel at a time: the emulated shared memory (g_smem, the kernels' `static` __shared__ variables) exists once, and the
// asynchronous entry points launch from several host threads
static std::mutex g_launch_mu;

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    if (smem > sizeof g_smem) abort();
    std::lock_guard<std::mutex> lk(g_launch_mu);
    const unsigned nwarps = (block + 31) / 32;
    // one set of OS threads per launch: thread t plays CUDA thread t of block 0, then of block 1, ... (the blocks run one after
    // another anyway - the shared memory exists once - and creating `block` threads per block dominated launches of many small blocks)
    BlockCtx ctx;
    ctx.block_bar.reset(new std::barrier<>(block));
    ctx.xchg.resize(nwarps);
    for (unsigned w = 0; w < nwarps; ++w) {
        unsigned n = std::min(32u, block - 32 * w);
        ctx.warp_bar.emplace_back(new std::barrier<>(n));
    }
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
        th.emplace_back([&, t] {
            g_ctx = &ctx;
            g_tid = uint3{t, 0, 0};
            g_bdim = dim3(block);
            g_gdim = dim3(grid);
            for (unsigned b = 0; b < grid; ++b) {
                g_bid = uint3{b, 0, 0};
                body();
                ctx.block_bar->arrive_and_wait();      // every thread has left block b before anyone touches shared memory as block b + 1
            }
        });
    for (auto& x : th) x.join();
}
}  // namespace nsb_emu

using namespace nsb;

static void make_tw(std::vector<f2>& tw) {
    tw.resize(kTwF2);
    for (int j = 1; j < 32; ++j)
        for (int l = 0; l < 32; ++l) {
            double a = -2.0 * M_PI * double(j * l) / 2048.0;
            tw[(j - 1) * 32 + l].x = float(std::cos(a));
            tw[(j - 1) * 32 + l].y = float(std::sin(a));
        }
}
struct Lane { c2 z[32]; };

extern "C" {
// lane-level checks of the frame transform: out = rfft(x), x = irfft(X) (numpy conventions)
void emu_rfft2048(const float* x, float* out_re, float* out_im) {
    std::vector<f2> tw; make_tw(tw);
    std::vector<f2> scratch(kScratchF2);
    Lane L[32];
    for (int lane = 0; lane < 32; ++lane)
        for (int t = 0; t < 32; ++t) L[lane].z[t] = mk2(0.5f * x[64 * t + lane], 0.5f * x[64 * t + 32 + lane]);
    for (int lane = 0; lane < 32; ++lane) fwd_phase1(L[lane].z, lane, scratch.data(), tw.data());
    for (int lane = 0; lane < 32; ++lane) fwd_phase2(L[lane].z, lane, scratch.data());
    for (int lane = 0; lane < 32; ++lane)
        for (int p = 0; p < 32; ++p) {
            if (lane == 0 && p == 0) { out_re[0] = L[0].z[0].x; out_im[0] = 0.f; out_re[1024] = L[0].z[0].y; out_im[1024] = 0.f; continue; }
            int k = bin_of(lane, p);
            out_re[k] = L[lane].z[p].x;
            out_im[k] = slot_is_conj(lane, p) ? -L[lane].z[p].y : L[lane].z[p].y;
        }
}
void emu_irfft2048(const float* in_re, const float* in_im, float* x) {
    std::vector<f2> tw; make_tw(tw);
    std::vector<f2> scratch(kScratchF2);
    Lane L[32];
    for (int lane = 0; lane < 32; ++lane)
        for (int p = 0; p < 32; ++p) {
            if (lane == 0 && p == 0) { L[0].z[0] = mk2(in_re[0], in_re[1024]); continue; }
            int k = bin_of(lane, p);
            L[lane].z[p] = mk2(in_re[k], slot_is_conj(lane, p) ? -in_im[k] : in_im[k]);
        }
    for (int lane = 0; lane < 32; ++lane) inv_phase1(L[lane].z, lane, scratch.data(), tw.data());
    for (int lane = 0; lane < 32; ++lane) inv_phase2(L[lane].z, lane, scratch.data());
    for (int lane = 0; lane < 32; ++lane)
        for (int t = 0; t < 32; ++t) {
            x[64 * t + lane] = L[lane].z[t].x * (1.0f / 2048.0f);
            x[64 * t + 32 + lane] = L[lane].z[t].y * (1.0f / 2048.0f);
        }
}
void emu_fft32(float* re, float* im, int dir) {
    c2 z[32];
    for (int i = 0; i < 32; ++i) z[i] = mk2(re[i], im[i]);
    if (dir < 0) fft32<-1>(z); else fft32<+1>(z);
    for (int i = 0; i < 32; ++i) { re[i] = z[i].x; im[i] = z[i].y; }
}
}  // extern "C"
