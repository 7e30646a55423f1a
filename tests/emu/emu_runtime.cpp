// TEST INFRASTRUCTURE ONLY: thread-per-CUDA-thread block executor for cuda_emu.h, plus lane-level
// entry points used by tests/test_emulated_kernels.py.  See cuda_emu.h.
#include "cuda_emu.h"
#include "../../nspeech_b200/csrc/frame_fft.cuh"

namespace nsb_emu {
thread_local BlockCtx* g_ctx = nullptr;
thread_local uint3 g_tid, g_bid;
thread_local dim3 g_bdim, g_gdim;
alignas(256) static unsigned char g_smem[256 * 1024];
unsigned char* dyn_smem() { return g_smem; }

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    if (smem > sizeof g_smem) abort();
    const unsigned nwarps = (block + 31) / 32;
    for (unsigned b = 0; b < grid; ++b) {
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
            th.emplace_back([&, t, b] {
                g_ctx = &ctx;
                g_tid = uint3{t, 0, 0};
                g_bid = uint3{b, 0, 0};
                g_bdim = dim3(block);
                g_gdim = dim3(grid);
                body();
            });
        for (auto& x : th) x.join();
    }
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
struct Lane { float re[32], im[32]; };

extern "C" {
// lane-level checks of the frame transform: out = rfft(x), x = irfft(X) (numpy conventions)
void emu_rfft2048(const float* x, float* out_re, float* out_im) {
    std::vector<f2> tw; make_tw(tw);
    std::vector<f2> scratch(kScratchF2);
    Lane L[32];
    for (int lane = 0; lane < 32; ++lane)
        for (int t = 0; t < 32; ++t) { L[lane].re[t] = 0.5f * x[64 * t + lane]; L[lane].im[t] = 0.5f * x[64 * t + 32 + lane]; }
    for (int lane = 0; lane < 32; ++lane) fwd_phase1(L[lane].re, L[lane].im, lane, scratch.data(), tw.data());
    for (int lane = 0; lane < 32; ++lane) fwd_phase2(L[lane].re, L[lane].im, lane, scratch.data());
    for (int lane = 0; lane < 32; ++lane)
        for (int p = 0; p < 32; ++p) {
            if (lane == 0 && p == 0) { out_re[0] = L[0].re[0]; out_im[0] = 0.f; out_re[1024] = L[0].im[0]; out_im[1024] = 0.f; continue; }
            int k = bin_of(lane, p);
            out_re[k] = L[lane].re[p];
            out_im[k] = slot_is_conj(lane, p) ? -L[lane].im[p] : L[lane].im[p];
        }
}
void emu_irfft2048(const float* in_re, const float* in_im, float* x) {
    std::vector<f2> tw; make_tw(tw);
    std::vector<f2> scratch(kScratchF2);
    Lane L[32];
    for (int lane = 0; lane < 32; ++lane)
        for (int p = 0; p < 32; ++p) {
            if (lane == 0 && p == 0) { L[0].re[0] = in_re[0]; L[0].im[0] = in_re[1024]; continue; }
            int k = bin_of(lane, p);
            L[lane].re[p] = in_re[k];
            L[lane].im[p] = slot_is_conj(lane, p) ? -in_im[k] : in_im[k];
        }
    for (int lane = 0; lane < 32; ++lane) inv_phase1(L[lane].re, L[lane].im, lane, scratch.data(), tw.data());
    for (int lane = 0; lane < 32; ++lane) inv_phase2(L[lane].re, L[lane].im, lane, scratch.data());
    for (int lane = 0; lane < 32; ++lane)
        for (int t = 0; t < 32; ++t) {
            x[64 * t + lane] = L[lane].re[t] * (1.0f / 2048.0f);
            x[64 * t + 32 + lane] = L[lane].im[t] * (1.0f / 2048.0f);
        }
}
void emu_fft32(float* re, float* im, int dir) {
    float r[32], i[32];
    memcpy(r, re, sizeof r); memcpy(i, im, sizeof i);
    if (dir < 0) fft32<-1>(r, i); else fft32<+1>(r, i);
    memcpy(re, r, sizeof r); memcpy(im, i, sizeof i);
}
}  // extern "C"
